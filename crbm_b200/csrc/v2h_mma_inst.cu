// One translation unit per number of MMA k-steps: compiled with -DCRBM_KS=<1..9> (motif_length <= 4*KS).
#include "v2h_mma.cuh"

#ifndef CRBM_KS
#error "compile with -DCRBM_KS=<1..9>"
#endif

namespace crbm {

template <int KS, int FLAGS, bool POOL>
static int launch_mma_pool(const V2HParams &p, dim3 grid, cudaStream_t stream) {
    static bool configured = false;
    constexpr size_t smem = mma_smem_bytes<KS, FLAGS>();
    if (!configured) {
        CRBM_CHECK_CUDA(cudaFuncSetAttribute(v2h_mma_kernel<KS, FLAGS, POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    CRBM_CHECK_CUDA(launch_chain(v2h_mma_kernel<KS, FLAGS, POOL>, grid, dim3(256), smem, stream, p));
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}
template <int KS, int FLAGS>
static int launch_mma_variant(const V2HParams &p, dim3 grid, cudaStream_t stream) {
    return p.pooling == 1 ? launch_mma_pool<KS, FLAGS, false>(p, grid, stream) : launch_mma_pool<KS, FLAGS, true>(p, grid, stream);
}

template <>
int launch_v2h_mma<CRBM_KS>(int flags, const V2HParams &p, dim3 grid, cudaStream_t stream) {
    switch (flags) {
#if CRBM_KS <= 4
        case F_STAT_P | F_STAT_Q: return launch_mma_variant<CRBM_KS, F_STAT_P | F_STAT_Q>(p, grid, stream);
#endif
        case F_STAT_P: return launch_mma_variant<CRBM_KS, F_STAT_P>(p, grid, stream);
        case F_STAT_Q: return launch_mma_variant<CRBM_KS, F_STAT_Q>(p, grid, stream);
        case F_STAT_P | F_SAMPLE: return launch_mma_variant<CRBM_KS, F_STAT_P | F_SAMPLE>(p, grid, stream);
        case F_SAMPLE: return launch_mma_variant<CRBM_KS, F_SAMPLE>(p, grid, stream);
        default:
            set_error("v2h_mma<KS=%d>: variant 0x%x not built", CRBM_KS, flags);
            return CRBM_ERR_UNSUPPORTED;
    }
}

template <>
int launch_v2h_mma_dual<CRBM_KS>(const V2HDualParams &d, dim3 grid, cudaStream_t stream) {
#if CRBM_KS <= 4
    static bool configured = false;
    constexpr size_t sa = mma_smem_bytes<CRBM_KS, F_STAT_P | F_STAT_Q>(), sb = mma_smem_bytes<CRBM_KS, F_STAT_P | F_SAMPLE>();
    constexpr size_t smem = sa > sb ? sa : sb;
    if (!configured) {
        CRBM_CHECK_CUDA(cudaFuncSetAttribute(v2h_mma_dual_kernel<CRBM_KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    CRBM_CHECK_CUDA(launch_chain(v2h_mma_dual_kernel<CRBM_KS>, grid, dim3(256), smem, stream, d));
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
#else
    (void)d; (void)grid; (void)stream;
    set_error("v2h_mma_dual<KS=%d>: the two-role kernel is built for motif_length <= 16", CRBM_KS);
    return CRBM_ERR_UNSUPPORTED;
#endif
}

template <>
int launch_v2h_dense_mma<CRBM_KS>(const DenseParams &p, dim3 grid, cudaStream_t stream) {
    static bool configured = false;
    constexpr size_t smem = MmaShape<CRBM_KS>::smem_bytes + 8 * MTILE * 34 * sizeof(float);
    if (!configured) {
        CRBM_CHECK_CUDA(cudaFuncSetAttribute(v2h_dense_mma_kernel<CRBM_KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    v2h_dense_mma_kernel<CRBM_KS><<<grid, 256, smem, stream>>>(p);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

template <>
int launch_v2h_scan_reduce<CRBM_KS>(const ScanReduceParams &sp, cudaStream_t stream) {
    static bool configured = false;
    constexpr size_t smem = MmaShape<CRBM_KS>::smem_bytes;
    if (!configured) {
        CRBM_CHECK_CUDA(cudaFuncSetAttribute(v2h_scan_reduce_kernel<CRBM_KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int mtiles = (sp.d.K + MTILE - 1) / MTILE;
    int gy = (sp.d.n_rows + 7) / 8;
    const int cap = (3 * num_sms()) / mtiles > 0 ? (3 * num_sms()) / mtiles : 1;
    if (gy > cap) gy = cap;
    v2h_scan_reduce_kernel<CRBM_KS><<<dim3(mtiles, gy), 256, smem, stream>>>(sp);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

}  // namespace crbm
