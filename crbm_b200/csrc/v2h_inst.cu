// One translation unit per padded motif length: compiled with -DCRBM_MW=<3,6,...,33> (see crbm_b200/build.py).
#include "v2h.cuh"

#ifndef CRBM_MW
#error "compile with -DCRBM_MW=<multiple of 3>"
#endif

namespace crbm {

template <int MW, int FLAGS>
static int launch_variant(const V2HParams &p, dim3 grid, cudaStream_t stream) {
    static bool configured = false;
    constexpr size_t smem = fused_smem_bytes<MW>();
    if (!configured) {
        CRBM_CHECK_CUDA(cudaFuncSetAttribute(v2h_fused_kernel<MW, FLAGS>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    v2h_fused_kernel<MW, FLAGS><<<grid, 256, smem, stream>>>(p);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

template <>
int launch_v2h_fused<CRBM_MW>(int flags, const V2HParams &p, dim3 grid, cudaStream_t stream) {
    switch (flags) {
#if CRBM_MW <= 15
        case F_STAT_P | F_STAT_Q: return launch_variant<CRBM_MW, F_STAT_P | F_STAT_Q>(p, grid, stream);
#endif
        case F_STAT_P: return launch_variant<CRBM_MW, F_STAT_P>(p, grid, stream);
        case F_STAT_Q: return launch_variant<CRBM_MW, F_STAT_Q>(p, grid, stream);
        case F_STAT_P | F_SAMPLE: return launch_variant<CRBM_MW, F_STAT_P | F_SAMPLE>(p, grid, stream);
        case F_SAMPLE: return launch_variant<CRBM_MW, F_SAMPLE>(p, grid, stream);
        case F_FE: return launch_variant<CRBM_MW, F_FE>(p, grid, stream);
        case F_FE | F_COUNT: return launch_variant<CRBM_MW, F_FE | F_COUNT>(p, grid, stream);
        default:
            set_error("v2h_fused<MW=%d>: variant 0x%x not built", CRBM_MW, flags);
            return CRBM_ERR_UNSUPPORTED;
    }
}

template <>
int launch_v2h_dense<CRBM_MW>(const DenseParams &p, dim3 grid, cudaStream_t stream) {
    static bool configured = false;
    constexpr size_t smem = dense_smem_bytes<CRBM_MW>();
    if (!configured) {
        CRBM_CHECK_CUDA(cudaFuncSetAttribute(v2h_dense_kernel<CRBM_MW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    v2h_dense_kernel<CRBM_MW><<<grid, 256, smem, stream>>>(p);
    CRBM_CHECK_LAUNCH();
    return CRBM_OK;
}

}  // namespace crbm
