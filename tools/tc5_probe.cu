// Hardware probe for the tcgen05 features the attention kernels rely on (run on the B200 box):
//   * where an M=64 cta_group::1 accumulator lands in TMEM (lane offset 0 and 16),
//   * MN-major A / B descriptors taken from chunk-major token tiles ([chunk][row][16 B]),
//   * N=16 MMAs.
// Every case builds operand byte images on the host with the layout formulas the kernels use, runs a list of MMAs,
// dumps TMEM and compares with a host product.   nvcc -gencode arch=compute_100a,code=sm_100a tools/tc5_probe.cu
#include "../www24-rat_b200/csrc/tc5.cuh"
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>

struct Op { uint32_t a_off, a_lbo, a_sbo, b_off, b_lbo, b_sbo, idesc, tcol, tlane, accum; };
struct Prog { int nops; Op ops[64]; };

__global__ void __launch_bounds__(128) k_probe(const unsigned char* img, int img_bytes, Prog p, float* dump, int ncols) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tbase;
    for (int i = threadIdx.x; i < img_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(img)[i];
    if (threadIdx.x == 0) { tc5::mbar_init(&bar, 1); tc5::fence_mbar_init(); }
    if (threadIdx.x < 32) tc5::tmem_alloc(&tbase, 512);
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // zero the dumped TMEM region
    for (int c = 0; c < ncols; c += 8) {
        float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        tc5::tmem_st8(tbase + ((warp * 32) << 16) + c, z);
    }
    tc5::tmem_st_wait();
    tc5::fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
        tc5::fence_after_sync();
        const uint32_t s0 = tc5::smem_u32(smem);
        for (int i = 0; i < p.nops; ++i) {
            const Op& o = p.ops[i];
            tc5::mma_f16(tbase + (o.tlane << 16) + o.tcol, tc5::smem_desc(s0 + o.a_off, o.a_lbo, o.a_sbo),
                         tc5::smem_desc(s0 + o.b_off, o.b_lbo, o.b_sbo), o.idesc, o.accum);
        }
        tc5::mma_commit(&bar);
    }
    tc5::mbar_wait(&bar, 0);
    tc5::fence_after_sync();
    for (int c = 0; c < ncols; c += 8) {
        float v[8];
        tc5::tmem_ld8(tbase + ((warp * 32) << 16) + c, v);
        tc5::tmem_ld_wait();
        for (int k = 0; k < 8; ++k) dump[(warp * 32 + lane) * ncols + c + k] = v[k];
    }
    tc5::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tc5::tmem_dealloc(tbase, 512);
}

static std::vector<unsigned char> g_img;
static void put(size_t off, float v) { __half h = __float2half(v); memcpy(&g_img[off], &h, 2); }
// chunk-major tile with ROWS rows: element (r, c)
static size_t toff(size_t base, int ROWS, int r, int c) { return base + ((size_t)(c / 8) * ROWS + r) * 16 + (c % 8) * 2; }

static std::vector<float> run(const Prog& p, int ncols) {
    unsigned char* dimg; float* ddump;
    cudaMalloc(&dimg, g_img.size()); cudaMalloc(&ddump, 128 * ncols * 4);
    cudaMemcpy(dimg, g_img.data(), g_img.size(), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k_probe<<<1, 128, g_img.size(), 0>>>(dimg, (int)g_img.size(), p, ddump, ncols);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    std::vector<float> d(128 * ncols);
    cudaMemcpy(d.data(), ddump, d.size() * 4, cudaMemcpyDeviceToHost);
    cudaFree(dimg); cudaFree(ddump);
    return d;
}
static float rnd() { return (float)((rand() % 9) - 4) * 0.25f; }
static int m64_lane(int m, int off) { return (m % 16) + 32 * (m / 16) + off; }

int main() {
    srand(1);
    // ---------------- case A: M=64 N=64 K=16, both K-major; lane offset 0 and 16
    for (int off = 0; off <= 16; off += 16) {
        g_img.assign(64 * 16 * 2 * 2, 0);
        const size_t A0 = 0, B0 = 64 * 16 * 2;
        for (int m = 0; m < 64; ++m) { put(toff(A0, 64, m, 0), (float)(m + 1)); put(toff(A0, 64, m, 1), 1.0f); }
        for (int n = 0; n < 64; ++n) { put(toff(B0, 64, n, 0), 1.0f); put(toff(B0, 64, n, 1), 128.0f * (n + 1)); }
        Prog p{}; p.nops = 1;
        p.ops[0] = Op{(uint32_t)A0, 64 * 16, 128, (uint32_t)B0, 64 * 16, 128, tc5::instr_desc(tc5::FMT_F16, 64, 64), 0, (uint32_t)off, 0};
        auto d = run(p, 64);
        int bad = 0, nz = 0;
        for (int l = 0; l < 128; ++l) for (int c = 0; c < 64; ++c) if (d[l * 64 + c] != 0.f) ++nz;
        for (int m = 0; m < 64; ++m) for (int n = 0; n < 64; ++n)
            if (d[m64_lane(m, off) * 64 + n] != (float)(m + 1) + 128.0f * (n + 1)) ++bad;
        printf("A(lane off %d): nonzero %d (expect 4096), mismatches under lane=(m%%16)+32*(m/16)+off: %d\n", off, nz, bad);
        if (bad) {
            for (int l = 0; l < 128; l += 1) { float v = d[l * 64]; if (v != 0.f) printf("  lane %d col0 -> m=%d\n", l, (int)v - 128 - 1); }
        }
    }
    // ---------------- case B/C: P [64x64] block tile (rows = query i, cols = key j), V / dO token tile [128 rows x 16]
    {
        const size_t P0 = 0, V0 = 64 * 64 * 2;
        g_img.assign(V0 + 128 * 16 * 2, 0);
        std::vector<float> P(64 * 64), V(128 * 16);
        for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) { P[i * 64 + j] = rnd(); put(toff(P0, 64, i, j), P[i * 64 + j]); }
        for (int r = 0; r < 128; ++r) for (int c = 0; c < 16; ++c) { V[r * 16 + c] = rnd(); put(toff(V0, 128, r, c), V[r * 16 + c]); }
        for (int half = 0; half < 2; ++half) {
            Prog p{};
            // O[i][d] = sum_j P[i][j] V[64*half + j][d] : A K-major (64 rows), B MN-major from rows 64*half.. of the token tile
            for (int j = 0; j < 4; ++j)
                p.ops[p.nops++] = Op{(uint32_t)(P0 + j * 2 * 64 * 16), 64 * 16, 128, (uint32_t)(V0 + (64 * half + 16 * j) * 16), 128, 128 * 16,
                                     tc5::instr_desc(tc5::FMT_F16, 64, 16, 0, 1), 0, 0, (uint32_t)(j > 0)};
            // dV[j][d] = sum_i P[i][j] dO[64*half + i][d] : A MN-major from the same P tile
            for (int j = 0; j < 4; ++j)
                p.ops[p.nops++] = Op{(uint32_t)(P0 + j * 256), 128, 64 * 16, (uint32_t)(V0 + (64 * half + 16 * j) * 16), 128, 128 * 16,
                                     tc5::instr_desc(tc5::FMT_F16, 64, 16, 1, 1), 16, 16, (uint32_t)(j > 0)};
            auto d = run(p, 32);
            double e1 = 0, e2 = 0;
            for (int i = 0; i < 64; ++i) for (int c = 0; c < 16; ++c) {
                double o = 0, dv = 0;
                for (int j = 0; j < 64; ++j) { o += P[i * 64 + j] * V[(64 * half + j) * 16 + c]; dv += P[j * 64 + i] * V[(64 * half + j) * 16 + c]; }
                e1 = fmax(e1, fabs(o - d[m64_lane(i, 0) * 32 + c]));
                e2 = fmax(e2, fabs(dv - d[m64_lane(i, 16) * 32 + 16 + c]));
            }
            printf("B(half %d): P.V (A K-major, B MN-major N=16) max err %.3g ; C: P^T.dO (A MN-major, lane off 16) max err %.3g\n", half, e1, e2);
        }
    }
    // ---------------- case D: weight gradient  G[f][c] = sum_t X[t][f] Y[t][c]  (X: [128 tok x 128 feat], Y: [128 tok x 48])
    {
        const size_t X0 = 0, Y0 = 128 * 128 * 2;
        g_img.assign(Y0 + 128 * 48 * 2, 0);
        std::vector<float> X(128 * 128), Y(128 * 48);
        for (int t = 0; t < 128; ++t) for (int f = 0; f < 128; ++f) { X[t * 128 + f] = rnd(); put(toff(X0, 128, t, f), X[t * 128 + f]); }
        for (int t = 0; t < 128; ++t) for (int c = 0; c < 48; ++c) { Y[t * 48 + c] = rnd(); put(toff(Y0, 128, t, c), Y[t * 48 + c]); }
        Prog p{};
        for (int j = 0; j < 8; ++j)
            p.ops[p.nops++] = Op{(uint32_t)(X0 + j * 256), 128, 128 * 16, (uint32_t)(Y0 + j * 256), 128, 128 * 16,
                                 tc5::instr_desc(tc5::FMT_F16, 128, 48, 1, 1), 0, 0, (uint32_t)(j > 0)};
        // second product on a feature sub-range: rows f = 64..127 only (M=64 from chunk 8 on), accumulated twice
        for (int rep = 0; rep < 2; ++rep)
            for (int j = 0; j < 8; ++j)
                p.ops[p.nops++] = Op{(uint32_t)(X0 + 8 * 128 * 16 + j * 256), 128, 128 * 16, (uint32_t)(Y0 + j * 256), 128, 128 * 16,
                                     tc5::instr_desc(tc5::FMT_F16, 64, 48, 1, 1), 48, 0, (uint32_t)(j > 0 || rep > 0)};
        auto d = run(p, 96);
        double e1 = 0, e2 = 0;
        for (int f = 0; f < 128; ++f) for (int c = 0; c < 48; ++c) {
            double g = 0;
            for (int t = 0; t < 128; ++t) g += X[t * 128 + f] * Y[t * 48 + c];
            e1 = fmax(e1, fabs(g - d[f * 96 + c]));
            if (f >= 64) e2 = fmax(e2, fabs(2 * g - d[m64_lane(f - 64, 0) * 96 + 48 + c]));
        }
        printf("D: wgrad X^T.Y (A, B MN-major from token tiles, K=128) max err %.3g ; M=64 sub-range accumulated twice max err %.3g\n", e1, e2);
    }
    // ---------------- case E: S = Q K^T per 64-row half of a 128-row token tile, both halves into the same columns
    {
        const size_t Q0 = 0, K0 = 128 * 16 * 2;
        g_img.assign(K0 + 128 * 16 * 2, 0);
        std::vector<float> Q(128 * 16), Kt(128 * 16);
        for (int r = 0; r < 128; ++r) for (int c = 0; c < 16; ++c) { Q[r * 16 + c] = rnd(); Kt[r * 16 + c] = rnd(); put(toff(Q0, 128, r, c), Q[r * 16 + c]); put(toff(K0, 128, r, c), Kt[r * 16 + c]); }
        Prog p{};
        for (int half = 0; half < 2; ++half)
            p.ops[p.nops++] = Op{(uint32_t)(Q0 + 64 * half * 16), 128 * 16, 128, (uint32_t)(K0 + 64 * half * 16), 128 * 16, 128,
                                 tc5::instr_desc(tc5::FMT_F16, 64, 64), 0, (uint32_t)(16 * half), 0};
        auto d = run(p, 64);
        double e = 0;
        for (int half = 0; half < 2; ++half) for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) {
            double s = 0;
            for (int c = 0; c < 16; ++c) s += Q[(64 * half + i) * 16 + c] * Kt[(64 * half + j) * 16 + c];
            e = fmax(e, fabs(s - d[m64_lane(i, 16 * half) * 64 + j]));
        }
        printf("E: S halves (A/B K-major sub-ranges of a 128-row tile, lane off 0/16) max err %.3g\n", e);
    }
    return 0;
}
