// Cross-GPU barrier for the tensor-parallel GEMM: one tiny kernel behind quanta_gemm_wna16_scatter.
//
// The fused gather needs the ranks to meet once per call: "every rank's tiles have landed in my buffer".  A generic
// barrier kernel costs a full launch boundary on each side.  This one is launched with programmatic stream
// serialization and releases ITS dependents at once: the next layer's GEMM grid may start while the barrier is
// still waiting for the peers — its weight stream does not depend on anything the barrier orders (common.cuh, pdl_wait) —
// and only that GEMM's activation loads and writes wait for the barrier to complete.  In a chain of column-parallel
// layers the NVLink round trip of call n hides behind the weight prefetch of call n + 1.
#include "common.cuh"

namespace quanta {

struct PeerBarrierArgs {
    unsigned int* flags[8];        // flags[r]: rank r's flag array (world unsigned ints, peer-mapped symmetric memory)
    unsigned int* epoch_counter;   // this rank's call counter (local device memory; shared with PeerSync)
    int rank, world;
};

__global__ void __launch_bounds__(32) peer_barrier_kernel(const __grid_constant__ PeerBarrierArgs a) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");       // the GEMM has completed; its stores, remote ones included, are flushed
    const int lane = threadIdx.x;
    unsigned int epoch = 0u;
    if (lane == 0) { epoch = *a.epoch_counter + 1u; *a.epoch_counter = epoch; }
    epoch = __shfl_sync(0xffffffffu, epoch, 0);
    __threadfence_system();
    if (lane < a.world && lane != a.rank) {
        // one lane per peer: all signals leave together, all peers are polled together
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.flags[lane] + a.rank), "r"(epoch) : "memory");
        unsigned int seen = 0, spins = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.flags[a.rank] + lane) : "memory");
            if ((int)(seen - epoch) >= 0 || ++spins > (1u << 28)) break;
        }
    }
    __syncwarp();
}

}  // namespace quanta

using namespace quanta;

extern "C" int quanta_peer_barrier(void* const* peer_flags, int rank, int world, unsigned int* epoch_counter, void* stream) {
    if (!peer_flags || !epoch_counter || world < 2 || world > 8 || rank < 0 || rank >= world) return QUANTA_EINVAL;
    PeerBarrierArgs a;
    for (int r = 0; r < 8; ++r) {
        a.flags[r] = r < world ? static_cast<unsigned int*>(peer_flags[r]) : nullptr;
        if (r < world && !a.flags[r]) return QUANTA_EINVAL;
    }
    a.epoch_counter = epoch_counter; a.rank = rank; a.world = world;
    cudaError_t e = launch_pdl(peer_barrier_kernel, dim3(1), dim3(32), 0, static_cast<cudaStream_t>(stream), a);
    return cuda_status(e != cudaSuccess ? e : cudaGetLastError());
}
