// attention backward (tcgen05 path), head width 20: one translation unit per head width keeps the build parallel.
#include "encoder_tc_attnbwd.cuh"

using namespace rat;

template <int DH, int KCH, bool VEC4, int JW>
static int launch_attn_bwd_tc(const AttnBwdTcArgs& a, int grid, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_attn_bwd_tc<DH, KCH, VEC4, JW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             max_smem_optin() - 2048);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_attn_bwd_tc)");
        attr_set = true;
    }
    k_attn_bwd_tc<DH, KCH, VEC4, JW><<<grid, TC_THREADS, a.smem_bytes, st>>>(a);
    RAT_CHECK_LAUNCH("k_attn_bwd_tc");
    return RAT_OK;
}
template <int DH, int KCH, bool VEC4>
static int launch_attn_bwd_tc_j(const AttnBwdTcArgs& a, int grid, cudaStream_t st) {
    const int jw = (a.njobs + 15) / 16;          // accumulator tiles per warp: instantiated for 3, 5 and 6
    if (jw <= 3) return launch_attn_bwd_tc<DH, KCH, VEC4, 3>(a, grid, st);
    if (jw <= 5) return launch_attn_bwd_tc<DH, KCH, VEC4, 5>(a, grid, st);
    if (jw <= 6) return launch_attn_bwd_tc<DH, KCH, VEC4, 6>(a, grid, st);
    return 1;
}
template <int DH>
static int launch_attn_bwd_tc_dh_impl(const AttnBwdTcArgs& a, int grid, cudaStream_t st) {
    const int kch = a.Kp / 16;
    const bool v4 = (a.D % 4) == 0;
#define RAT_AB(K_) (v4 ? launch_attn_bwd_tc_j<DH, K_, true>(a, grid, st) : launch_attn_bwd_tc_j<DH, K_, false>(a, grid, st))
    switch (kch) {
        case 1: return RAT_AB(1);
        case 2: return RAT_AB(2);
        case 3: return RAT_AB(3);
        default: return 1;
    }
#undef RAT_AB
}


int attn_bwd_tc_launch_dh20(const AttnBwdTcArgs& a, int grid, cudaStream_t st) { return launch_attn_bwd_tc_dh_impl<20>(a, grid, st); }
