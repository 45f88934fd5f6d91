// fp32 FFMA GEMM with multi-segment K (split-K over operand sources) and fused bias /
// per-caption row-add / element-wise add epilogue.  C[M][N] = sum_s A_s[M][K_s] * W[N][K_s]^T.
//
// This is the exact-fp32 path: it serves the accuracy-critical projections (att_va prologue,
// attention-score inputs) and is the verification twin of the tcgen05 kernels.
// Replaces the cuBLAS SGEMMs behind nn.Linear / nn.LSTMCell at
// controllable_captioning.py:151-152,155,161-163,177-178,181,184 (SURVEY.md §2.1 k5-k7,k9,k10).
#include "common.cuh"

namespace vsr {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, THREADS = 256;

__global__ void __launch_bounds__(THREADS) k_gemm_simt(const GemmArgs g) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Ws[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  if (g.row_skip != nullptr) {  // skip tiles whose rows are all padding (block-uniform)
    int any = 0;
    if (tid < BM && m0 + tid < g.M) any = g.row_skip[m0 + tid];
    if (!__syncthreads_or(any)) return;
  }

  const int lrow = tid >> 2;        // 0..63 : tile row loaded by this thread
  const int lk = (tid & 3) << 2;    // 0,4,8,12 : k offset of its float4
  const int tx = tid & 15, ty = tid >> 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // flattened iteration over (segment, k-tile)
  int total_tiles = 0;
  for (int s = 0; s < g.nseg; ++s) total_tiles += g.seg[s].k / BK;

  auto load_tile = [&](int tile, float4& av, float4& wv) {
    int s = 0, kt = tile, wk = 0;
    while (kt >= g.seg[s].k / BK) { kt -= g.seg[s].k / BK; wk += g.seg[s].k; ++s; }
    const GemmSeg& sg = g.seg[s];
    const int k = kt * BK + lk;
    const int row = m0 + lrow;
    av = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < g.M && k + 4 <= sg.k_valid)
      av = *reinterpret_cast<const float4*>(sg.a + (size_t)row * sg.lda + k);
    wv = *reinterpret_cast<const float4*>(g.w + (size_t)(n0 + lrow) * g.ldw + wk + k);
  };
  auto store_tile = [&](int buf, const float4& av, const float4& wv) {
    As[buf][lk + 0][lrow] = av.x; As[buf][lk + 1][lrow] = av.y;
    As[buf][lk + 2][lrow] = av.z; As[buf][lk + 3][lrow] = av.w;
    Ws[buf][lk + 0][lrow] = wv.x; Ws[buf][lk + 1][lrow] = wv.y;
    Ws[buf][lk + 2][lrow] = wv.z; Ws[buf][lk + 3][lrow] = wv.w;
  };

  float4 av, wv;
  load_tile(0, av, wv);
  store_tile(0, av, wv);
  __syncthreads();
  for (int tile = 0; tile < total_tiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < total_tiles) load_tile(tile + 1, av, wv);  // global prefetch overlaps the FMAs
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float ar[4] = {a.x, a.y, a.z, a.w};
      const float wr[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
    if (tile + 1 < total_tiles) store_tile(buf ^ 1, av, wv);
    __syncthreads();
  }

  const int n = n0 + tx * 4;
  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (g.bias != nullptr) bv = *reinterpret_cast<const float4*>(g.bias + n);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= g.M) continue;
    float4 o = make_float4(acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w);
    if (g.rowadd != nullptr) {
      const float4 r = *reinterpret_cast<const float4*>(
          g.rowadd + (size_t)((row / g.row_div) * g.rowadd_mul) * g.ld_rowadd + n);
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    if (g.cadd != nullptr) {
      const float4 r = *reinterpret_cast<const float4*>(g.cadd + (size_t)row * g.ld_cadd + n);
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    if (g.gather != nullptr) {
      const float4 r = *reinterpret_cast<const float4*>(g.gather + (size_t)g.gather_idx[row] * g.ld_gather + n);
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    if (g.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    *reinterpret_cast<float4*>(g.c + (size_t)row * g.ldc + n) = o;
  }
}

}  // namespace

bool gemm_uses_tc(const Ctx* c, const GemmArgs& g) {
  bool tc = c->use_tc && g.wb != nullptr && g.wb->hi != nullptr;
  for (int s = 0; s < g.nseg && tc; ++s) tc = g.seg[s].b != nullptr && g.seg[s].b->hi != nullptr;
  return tc;
}

int launch_gemm(Ctx* c, const GemmArgs& g, cudaStream_t st, const GemmArgs* g2) {
  if (gemm_uses_tc(c, g) && (g2 == nullptr || gemm_uses_tc(c, *g2))) return launch_gemm_tc(g, g2, st);
  VSR_TRY(launch_gemm_simt(g, st));
  if (g2 != nullptr) VSR_TRY(launch_gemm_simt(*g2, st));
  return VSR_OK;
}

int launch_gemm_simt(const GemmArgs& g, cudaStream_t st) {
  VSR_REQUIRE(g.N % BN == 0 && g.M > 0 && g.nseg >= 1 && g.nseg <= 3, VSR_EINVAL,
              "launch_gemm: bad shape M=%d N=%d nseg=%d", g.M, g.N, g.nseg);
  for (int s = 0; s < g.nseg; ++s)
    VSR_REQUIRE(g.seg[s].k % BK == 0 && g.seg[s].k_valid % 4 == 0 && g.seg[s].lda % 4 == 0,
                VSR_EINVAL, "launch_gemm: segment %d k=%d k_valid=%d lda=%d not aligned", s,
                g.seg[s].k, g.seg[s].k_valid, g.seg[s].lda);
  dim3 grid(g.N / BN, (g.M + BM - 1) / BM);
  k_gemm_simt<<<grid, THREADS, 0, st>>>(g);
  VSR_CHECK_CUDA(cudaGetLastError());
  return VSR_OK;
}

}  // namespace vsr
