// attention backward (tcgen05 path): planning, workspace, dispatch and the fixed-order reduction of the per-CTA records.
#include "encoder_tc_attnbwd.cuh"

using namespace rat;

int attn_bwd_tc_launch_dh8(const AttnBwdTcArgs& a, int grid, cudaStream_t st);
int attn_bwd_tc_launch_dh10(const AttnBwdTcArgs& a, int grid, cudaStream_t st);
int attn_bwd_tc_launch_dh20(const AttnBwdTcArgs& a, int grid, cudaStream_t st);

// ---- attention backward (tcgen05) host side --------------------------------------------------------------------
static bool attn_bwd_tc_plan(int S, int D, int heads, int dh, AttnBwdTcArgs* a) {
    if (S > 16 || S < 1 || (dh != 10 && dh != 20 && dh != 8) || D < 2 || (D & 1) || D > 64) return false;
    const int DHP = pad16(dh);
    a->Kp = pad16(D);
    int hc = 0;
    for (int c = heads; c >= 1; --c) {
        if (heads % c) continue;
        const int ncq = 3 * c * DHP, ndo = c * DHP;
        if (ncq <= 256 && ncq + ndo + a->Kp <= 512) { hc = c; break; }
    }
    if (!hc) return false;
    a->D = D; a->H = heads; a->I = heads * dh;
    a->hc = hc; a->nchunks = heads / hc;
    a->NCq = 3 * hc * DHP; a->NCc = pad16(3 * hc * dh); a->NDo = hc * DHP; a->Cc = pad16(hc * dh);
    if (a->NCc > 256) return false;
    a->SPT = TILE_M / S;
    if (S <= 8) a->SPT &= ~1;
    {   // The attention core runs ceil(tasks / 16 warps) rounds per head chunk; a slightly smaller tile can save a whole
        // round (S = 14: 9 sequences x 4 heads = 36 tasks = 3 rounds, 8 sequences = 32 tasks = 2 rounds).  Cost model
        // per tile: fixed phases ~ 7.3 round-equivalents (measured split 55 % / 45 % at 6 rounds) + rounds.
        const int step = S <= 8 ? 2 : 1;
        double best = 1e30;
        int best_spt = a->SPT;
        for (int spt = a->SPT; spt >= std::max(step, a->SPT - 3 * step); spt -= step) {
            const int tasks = (S <= 8 ? spt / 2 : spt) * hc;
            const double cost = (7.3 + (double)a->nchunks * ((tasks + 15) / 16)) / ((double)spt * S);
            if (cost < best - 1e-9) { best = cost; best_spt = spt; }
        }
        a->SPT = best_spt;
    }
    const int NP = a->Kp / 16;
    a->njobs = a->nchunks * ((a->NCc / 16 + a->Cc / 16) * NP) + 3 * NP;
    if ((a->njobs + 15) / 16 > 6) return false;
    {   // TMEM: q|k|v + dO + dA accumulators + the parking columns of the weight-gradient accumulators (4 warps per
        // lane quadrant x JW x 8; JW as instantiated: 3, 5 or 6)
        const int jw = (a->njobs + 15) / 16, jwi = jw <= 3 ? 3 : jw <= 5 ? 5 : 6;
        if (a->NCq + a->NDo + a->Kp + 4 * jwi * 8 > 512) return false;
    }
    a->psize = a->nchunks * (a->NCc + a->Cc) * a->Kp + 3 * a->Kp;
    const size_t images = (size_t)a->nchunks * (a->NCq + a->NCc + a->NDo) * a->Kp * 2;
    const size_t tiles = (size_t)TILE_M * (2 * a->Kp + a->NCq + a->NCc + a->Cc + a->NDo) * 2;
    if ((size_t)2 * TILE_M * a->Kp * 2 > (size_t)TILE_M * a->NCq * 2) return false;      // G1|G2 alias the q|k|v tile
    a->smem_bytes = (int)(images + tiles + (size_t)(2 + 8) * TILE_M * 4);
    return a->smem_bytes <= max_smem_optin() - 2048;
}
static int attn_bwd_tc_grid(const AttnBwdTcArgs& a, long long nseq) {
    return (int)std::min<long long>((nseq + a.SPT - 1) / a.SPT, (long long)num_sms());
}
size_t attn_bwd_tc_workspace_bytes(int B, int T, int N, int D, int heads, int dh, int mode) {
    AttnBwdTcArgs a{};
    const int S = mode == 0 ? N : T;
    if (!attn_bwd_tc_plan(S, D, heads, dh, &a)) return 0;
    const long long nseq = mode == 0 ? (long long)B * T : (long long)B * N;
    return (size_t)attn_bwd_tc_grid(a, nseq) * a.psize * sizeof(float);
}

int attn_bwd_tc_dispatch(const float* x, const float* dout, const float* base, float* dx, const float* ln_w,
                         const float* ln_b, const float* Wq, const float* Wk, const float* Wv, const float* Wo, float* dWq,
                         float* dWk, float* dWv, float* dWo, float* dbo, float* dln_w, float* dln_b, int accumulate_wq,
                         int B, int T, int N, int D, int heads, int dh, float scale, float alpha, int mode,
                         const float* dout_amax, float* dx_amax, float* workspace, size_t workspace_bytes, float out_drop_p,
                         unsigned long long seed, unsigned int rng_stream, cudaStream_t st) {
    AttnBwdTcArgs a{};
    const int S = mode == 0 ? N : T;
    if (!attn_bwd_tc_plan(S, D, heads, dh, &a)) return 1;
    a.nseq = mode == 0 ? (long long)B * T : (long long)B * N;
    const int grid = attn_bwd_tc_grid(a, a.nseq);
    if (!workspace || workspace_bytes < (size_t)grid * a.psize * sizeof(float)) return 1;
    reduce_ws_acquire(st, workspace);       // a deferred reduction may still be reading the records of an earlier call
    a.x = x; a.dout = dout; a.base = base; a.dx = dx; a.ln_w = ln_w; a.ln_b = ln_b; a.Wq = Wq; a.Wk = Wk; a.Wv = Wv; a.Wo = Wo;
    a.partials = workspace; a.dout_amax = dout_amax; a.dx_amax = dx_amax;
    a.g.S = S; a.g.mode = mode; a.g.T = T; a.g.N = N;
    a.scale = scale; a.alpha = alpha;
    a.out_drop_p = out_drop_p; a.seed = seed; a.rng_stream = rng_stream; a.rng_step = rng_step_ptr();
    int rc;
    switch (dh) {
        case 8: rc = attn_bwd_tc_launch_dh8(a, grid, st); break;
        case 10: rc = attn_bwd_tc_launch_dh10(a, grid, st); break;
        case 20: rc = attn_bwd_tc_launch_dh20(a, grid, st); break;
        default: return 1;
    }
    if (rc != RAT_OK) return rc;
    AttnReduceTcArgs r{workspace, grid, a.psize, dWq, dWk, dWv, dWo, dbo, dln_w, dln_b, accumulate_wq, D, a.I, dh, a.hc,
                       a.nchunks, a.Kp, a.NCc, a.Cc, 1.0f, 1.0f};
    const int total = 4 * a.I * D + 3 * D;
    cudaStream_t rs = reduce_fork(st, workspace);
    k_reduce_attn_tc<<<std::max(1, std::min((total + 31) / 32, 1024)), dim3(32, 8), 0, rs>>>(r);
    RAT_CHECK_LAUNCH("k_reduce_attn_tc");
    reduce_forked(rs, st, workspace);
    return RAT_OK;
}

