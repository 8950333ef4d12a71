// One translation unit per pooling factor of the rows-on-TMEM-lanes scan / free-energy kernel: compiled with -DCRBM_TT_P=<1,2,4,5,10>.
#include "scan_rt.cuh"
#include "scan_rd.cuh"

#ifndef CRBM_TT_P
#error "compile with -DCRBM_TT_P=<pooling factor>"
#endif

namespace crbm {

template <int MODE>
static int launch_rt_variant(const RowsTcParams &p, dim3 grid, cudaStream_t stream) {
    static size_t configured[64];                         // per device
    int dev = 0;
    CRBM_CHECK_CUDA(cudaGetDevice(&dev));
    // the CTA allocates all 512 TMEM columns: ask for more than half of the shared memory so that two CTAs never share an SM
    size_t smem = rt_smem_bytes(p.n_atoms, p.xstages);
    if (smem < 120 * 1024) smem = 120 * 1024;
    if (smem > configured[dev & 63]) {
        CRBM_CHECK_CUDA(cudaFuncSetAttribute(v2h_rows_tc_kernel<MODE, CRBM_TT_P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev & 63] = smem;
    }
    v2h_rows_tc_kernel<MODE, CRBM_TT_P><<<grid, RT_THREADS, smem, stream>>>(p);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

template <>
int launch_v2h_rows_tc<CRBM_TT_P>(int mode, const RowsTcParams &p, dim3 grid, cudaStream_t stream) {
    if (mode == RT_FE) return launch_rt_variant<RT_FE>(p, grid, stream);
    // un-pooled scans that do not ask for the mean run on the activities alone (the pooled maximum needs the probabilities anyway)
#if CRBM_TT_P == 1
    if (!p.out_mean) return launch_rt_variant<RT_REDUCE_NOMEAN>(p, grid, stream);
#endif
    return launch_rt_variant<RT_REDUCE>(p, grid, stream);
}

template <>
int launch_v2h_rowdense_tc<CRBM_TT_P>(const RowDenseParams &p, const CUtensorMap &tmap, dim3 grid, cudaStream_t stream) {
    static size_t configured[64];                         // per device
    int dev = 0;
    CRBM_CHECK_CUDA(cudaGetDevice(&dev));
    size_t smem = rd_smem_bytes(p.n_atoms, p.xstages);
    if (smem < 120 * 1024) smem = 120 * 1024;             // all 512 TMEM columns: one CTA per SM
    if (smem > configured[dev & 63]) {
        CRBM_CHECK_CUDA(cudaFuncSetAttribute(v2h_rowdense_tc_kernel<CRBM_TT_P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev & 63] = smem;
    }
    v2h_rowdense_tc_kernel<CRBM_TT_P><<<grid, RD_THREADS, smem, stream>>>(p, tmap);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

}  // namespace crbm
