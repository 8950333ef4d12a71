// One translation unit per pooling factor of the tcgen05 training kernel: compiled with -DCRBM_TT_P=<1,2,4,5,10>.
#include "train_tc.cuh"

#ifndef CRBM_TT_P
#error "compile with -DCRBM_TT_P=<pooling factor>"
#endif

namespace crbm {

template <int FLAGS>
static int launch_tt_variant(const TrainTcParams &p, dim3 grid, cudaStream_t stream) {
    static size_t configured[64];                         // per device (the attribute is per device)
    int dev = 0;
    CRBM_CHECK_CUDA(cudaGetDevice(&dev));
    // the CTA allocates all 512 TMEM columns: ask for more than half of the shared memory so that two CTAs never share an SM
    size_t smem = tt_smem_bytes(p.n_atoms, p.xstages, (FLAGS & (F_STAT_P | F_STAT_Q)) ? p.npsets : 0);
    if (smem < 120 * 1024) smem = 120 * 1024;
    if (smem > configured[dev & 63]) {
        CRBM_CHECK_CUDA(cudaFuncSetAttribute(v2h_train_tc_kernel<FLAGS, CRBM_TT_P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev & 63] = smem;
    }
    CRBM_CHECK_CUDA(launch_chain(v2h_train_tc_kernel<FLAGS, CRBM_TT_P>, grid, dim3(TT_THREADS), smem, stream, p));
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

template <>
int launch_v2h_train_tc<CRBM_TT_P>(int flags, const TrainTcParams &p, dim3 grid, cudaStream_t stream) {
    switch (flags) {
        case F_STAT_P | F_STAT_Q: return launch_tt_variant<F_STAT_P | F_STAT_Q>(p, grid, stream);
        case F_STAT_P | F_SAMPLE: return launch_tt_variant<F_STAT_P | F_SAMPLE>(p, grid, stream);
        case F_SAMPLE: return launch_tt_variant<F_SAMPLE>(p, grid, stream);
        default:
            set_error("v2h_train_tc<P=%d>: variant 0x%x not built", CRBM_TT_P, flags);
            return CRBM_ERR_UNSUPPORTED;
    }
}

}  // namespace crbm
