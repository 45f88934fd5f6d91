// S-level SSP of the eval pre-step: the semantic-role sorter S_SSP.generate(mode='not-normal'),
// batched over independent (verb, role set) problems, with no host round trip between the decoder steps.
//
//   reference                                                    here
//   TransformerEncoder.forward      sort_modules.py:52-63         k_sort_embed_enc, run_layers(encoder)
//   TransformerEncoderLayer         transformer_modules.py:325-346  k_sort_ln / GEMM / k_sort_attn / GEMM(+x) / k_sort_ln / GEMM(relu) / GEMM(+x)
//   MultiHeadAttention              transformer_modules.py:104-134  packed [Q;K;V] projection + k_sort_attn (one warp per head)
//   TransformerDecoder(+Layer)      sort_modules.py:79-99, 120-135  one new position per step against a key/value cache: position i >= 1
//                                                                 sees positions 1..i (the <bos> key is masked: weight exp(-1e3) = 0), position 0
//                                                                 sees only itself, so earlier positions never change when the prefix grows
//   S_SSP.generate 'not-normal'     sort_model.py:149-183         k_sort_select: log-softmax over the role ids, first maximum among the
//                                                                 roles still to be placed, in slot order
// The reference runs this per (caption, verb) with batch 1, re-running the decoder over the whole prefix at every step
// (eval_coco.py:170-174).  The projections (M = 10 P rows in the encoder, P rows per decoder step; 0.45 GFLOP per problem) run
// on the tcgen05 GEMM kernels of the decoder path in f16x3 mode (gemm_tc.cu: fp16 value + fp16 residual operands, three MMAs
// into one fp32 TMEM accumulator; fused bias / residual epilogue): the layer-norm, attention and ReLU kernels write the
// operand twins of what they produce.  The two tiny projections (fc_feat, the 512 -> 26 role head) and everything else are
// fp32 warp-level kernels.  VSRDEC_GEMM=simt switches every projection to the fp32 FFMA twin.
#include "common.cuh"
#include "twin_util.cuh"

namespace vsr {
namespace {

constexpr int SORT_MAXL = 4;        // layers per stack
constexpr int SORT_MAXK = 16;       // keys an attention query can see (max_len + 1 <= 16)
constexpr int SORT_DPL = 16;        // d_model / 32 values per lane in the head kernel (d_model <= 512)
constexpr int SORT_HD4 = 16;        // head dimension / 4 that fits the attention kernel's registers (head_dim <= 64)
constexpr int SORT_NPAD = 64;       // N padding of the FFMA GEMM

struct SortLayer {
  float *qkv_w, *qkv_b, *o_w, *o_b, *w1_w, *w1_b, *w2_w, *w2_b, *ln_w[3], *ln_b[3];
  // tensor-core twins of the weights; q_v / kv_v are row-range views of qkv_p ([0, d) and [d, 3d)) sharing its arrays and scale
  F16Pair qkv_p, q_v, kv_v, o_p, w1_p, w2_p;
};

struct SortCtx {
  int device = 0;
  VsrSortDims d{};
  float* wbuf = nullptr;            // one allocation holding every packed weight
  float *sr_emb = nullptr, *v_emb = nullptr, *fc_w = nullptr, *fc_b = nullptr;
  float *enc_ln_w = nullptr, *enc_ln_b = nullptr, *dec_ln_w = nullptr, *dec_ln_b = nullptr, *exp_w = nullptr, *exp_b = nullptr;
  SortLayer enc[SORT_MAXL], dec[SORT_MAXL];
  // workspace for cap_P problems
  int cap_P = 0;
  float* ws = nullptr;
  float *x = nullptr, *y = nullptr, *qkv = nullptr, *ctx = nullptr, *ff = nullptr, *prior = nullptr, *logits = nullptr;
  float *ckv[SORT_MAXL] = {}, *kc[SORT_MAXL] = {}, *vc[SORT_MAXL] = {};
  int32_t* token = nullptr;
  uint32_t* remain = nullptr;
  bool use_tc = true;
  F16Pair y_b, ctx_b, ff_b, prior_b;        // operand twins of the activations that feed a projection
};

// ---------------------------------------------------------------- kernels
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// remain[p] = bit i set iff roles[p][i] != 0 (sort_model.py:112); token[p] = <bos> = 0
__global__ void k_sort_init(const int64_t* __restrict__ roles, int P, int L, uint32_t* __restrict__ remain, int32_t* __restrict__ token) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  uint32_t m = 0;
  for (int i = 0; i < L; ++i) m |= (roles[(size_t)p * L + i] != 0 ? 1u : 0u) << i;
  remain[p] = m;
  token[p] = 0;
}

// x[p*L + j] = v_emb[verb_p] * sqrt(d) + sr_emb[role_pj] * sqrt(d)        sort_modules.py:54, transformer_modules.py:199-200
__global__ void k_sort_embed_enc(const int64_t* __restrict__ verbs, const int64_t* __restrict__ roles, const float* __restrict__ v_emb,
                                 const float* __restrict__ sr_emb, int L, int d, int n_verbs, int n_roles, float scale,
                                 float* __restrict__ x) {
  const int row = blockIdx.x, p = row / L;
  int64_t v = verbs[p], r = roles[row];
  v = v < 0 ? 0 : (v >= n_verbs ? n_verbs - 1 : v);
  r = r < 0 ? 0 : (r >= n_roles ? n_roles - 1 : r);
  const float4* ve = reinterpret_cast<const float4*>(v_emb + (size_t)v * d);
  const float4* re = reinterpret_cast<const float4*>(sr_emb + (size_t)r * d);
  float4* o = reinterpret_cast<float4*>(x + (size_t)row * d);
  for (int c = threadIdx.x; c < d / 4; c += blockDim.x) {
    const float4 a = ve[c], b = re[c];
    o[c] = make_float4(a.x * scale + b.x * scale, a.y * scale + b.y * scale, a.z * scale + b.z * scale, a.w * scale + b.w * scale);
  }
}

// x[p] = sr_emb[token_p] * sqrt(d)                                         sort_modules.py:126
__global__ void k_sort_embed_dec(const int32_t* __restrict__ token, const float* __restrict__ sr_emb, int d, float scale,
                                 float* __restrict__ x) {
  const int p = blockIdx.x;
  const float4* re = reinterpret_cast<const float4*>(sr_emb + (size_t)token[p] * d);
  float4* o = reinterpret_cast<float4*>(x + (size_t)p * d);
  for (int c = threadIdx.x; c < d / 4; c += blockDim.x) {
    const float4 b = re[c];
    o[c] = make_float4(b.x * scale, b.y * scale, b.z * scale, b.w * scale);
  }
}

// LayerNorm over the last dimension (eps 1e-5, biased variance), one warp per row; y and / or its operand twins
__global__ void k_sort_ln(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, int rows, int d,
                          float* __restrict__ y, const TwinOut tw) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * d;
  float s = 0.f;
  for (int c = lane; c < d; c += 32) s += xr[c];
  const float mean = wsum(s) / (float)d;
  float q = 0.f;
  for (int c = lane; c < d; c += 32) { const float t = xr[c] - mean; q += t * t; }
  const float rstd = rsqrtf(wsum(q) / (float)d + 1e-5f);
  for (int c = lane; c < d; c += 32) {
    const float v = (xr[c] - mean) * rstd * w[c] + b[c];
    if (y != nullptr) y[(size_t)row * d + c] = v;
    put_twin(tw, (size_t)row * d + c, v);
  }
}

// Multi-head attention of nq queries per problem over the keys [k0, k1) of that problem; one warp per head
// (transformer_modules.py:36-54: logits / sqrt(head_dim), softmax, weighted values).  With `store_pos >= 0` the problem's new key /
// value row (columns [d, 2d) and [2d, 3d) of its q row) is first appended to the cache at that position.
struct SortAttn {
  const float* q; int ldq; int nq;          // query row of (problem p, i): q + (p * nq + i) * ldq, head h at column h * hd
  const float* k; const float* v; int ldkv; // key row of (problem p, j): k + (p * kv_rows + j) * ldkv
  int kv_rows; int k0, k1;
  float* kc; float* vc; int store_pos;      // cache append (decoder self-attention) or store_pos < 0
  float* out; int ldo;                      // context row of (p, i): out + (p * nq + i) * ldo (fp32, or null)
  TwinOut out_tw;                           //   ... and / or its operand twins, same layout
  int hd; float inv_sqrt_hd;
};
// lane j owns key j (all hd values in registers, loaded once per problem and head); per query: hd FMAs per lane, two warp
// reductions for the softmax, then lanes run along the head dimension for the weighted values (weights broadcast by shuffle)
__global__ void k_sort_attn(const SortAttn a) {
  const int p = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hd = a.hd, col0 = h * hd, hd4 = hd >> 2;
  const int d = hd * (blockDim.x >> 5);
  const int nk = a.k1 - a.k0;
  const float* fresh = nullptr;         // decoder self-attention: the q row also carries the new key / value of position store_pos
  int jnew = -1;
  if (a.store_pos >= 0) {
    fresh = a.q + (size_t)p * a.ldq;
    jnew = a.store_pos - a.k0;
    float* kd = a.kc + ((size_t)p * a.kv_rows + a.store_pos) * a.ldkv;
    float* vd = a.vc + ((size_t)p * a.kv_rows + a.store_pos) * a.ldkv;
    for (int e = lane; e < hd; e += 32) { kd[col0 + e] = fresh[d + col0 + e]; vd[col0 + e] = fresh[2 * d + col0 + e]; }
  }
  float4 kreg[SORT_HD4];
  if (lane < nk) {
    const float* kr = lane == jnew ? fresh + d + col0 : a.k + ((size_t)p * a.kv_rows + a.k0 + lane) * a.ldkv + col0;
#pragma unroll
    for (int c = 0; c < SORT_HD4; ++c) kreg[c] = c < hd4 ? *reinterpret_cast<const float4*>(kr + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = 0; i < a.nq; ++i) {
    const float* qr = a.q + ((size_t)p * a.nq + i) * a.ldq + col0;
    float logit = -INFINITY;
    if (lane < nk) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < SORT_HD4; ++c) {
        if (c < hd4) {
          const float4 qv = *reinterpret_cast<const float4*>(qr + 4 * c);       // same address in every lane: one broadcast
          s = fmaf(qv.x, kreg[c].x, fmaf(qv.y, kreg[c].y, fmaf(qv.z, kreg[c].z, fmaf(qv.w, kreg[c].w, s))));
        }
      }
      logit = s * a.inv_sqrt_hd;
    }
    const float m = wmax(logit);
    const float ex = lane < nk ? expf(logit - m) : 0.f;
    const float wgt = ex / wsum(ex);
    const size_t o_off = ((size_t)p * a.nq + i) * a.ldo + col0;
    for (int e = lane; e < hd; e += 32) {
      float acc = 0.f;
      for (int j = 0; j < nk; ++j) {
        const float* vr = j == jnew ? fresh + 2 * d + col0 : a.v + ((size_t)p * a.kv_rows + a.k0 + j) * a.ldkv + col0;
        acc = fmaf(__shfl_sync(0xffffffffu, wgt, j), vr[e], acc);
      }
      if (a.out != nullptr) a.out[o_off + e] = acc;
      put_twin(a.out_tw, o_off + e, acc);
    }
  }
}

// log-softmax over the role ids of the last position and the constrained greedy choice (sort_model.py:158-180): the first
// maximum among the roles still to be placed, scanned in slot order; one warp per problem
// The whole head of a decoder step in the same warp: final layer norm of the new position's state (sort_modules.py:134), the
// d -> n_roles projection (sort_model.py:159; each role's weight row read coalesced, one warp reduction per role), then the choice.
__global__ void k_sort_select(const float* __restrict__ x, int d, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                              const float* __restrict__ exp_w, const float* __restrict__ exp_b, int n_roles,
                              const int64_t* __restrict__ roles, int L, int P,
                              int t, int n_steps, uint32_t* __restrict__ remain, int32_t* __restrict__ token, int64_t* __restrict__ pred,
                              float* __restrict__ logp, float* __restrict__ rows) {
  extern __shared__ __align__(16) float s_w[];           // [n_roles][d]: the projection, staged once per block
  for (int i = threadIdx.x * 4; i < n_roles * d; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(s_w + i) = __ldg(reinterpret_cast<const float4*>(exp_w + i));
  __syncthreads();
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (p >= P) return;
  const float* xr = x + (size_t)p * d;
  float yv[SORT_DPL];                                    // this lane's normalised values: columns lane, lane + 32, ...
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < SORT_DPL; ++i) { yv[i] = lane + 32 * i < d ? xr[lane + 32 * i] : 0.f; s += yv[i]; }
  const float mean = wsum(s) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < SORT_DPL; ++i) { const float u = lane + 32 * i < d ? yv[i] - mean : 0.f; q += u * u; }
  const float rstd = rsqrtf(wsum(q) / (float)d + 1e-5f);
#pragma unroll
  for (int i = 0; i < SORT_DPL; ++i) {
    const int c = lane + 32 * i;
    yv[i] = c < d ? (yv[i] - mean) * rstd * ln_w[c] + ln_b[c] : 0.f;
  }
  float v = -INFINITY;
  for (int r = 0; r < n_roles; ++r) {
    const float* w = s_w + r * d;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < SORT_DPL; ++i) if (lane + 32 * i < d) acc = fmaf(yv[i], w[lane + 32 * i], acc);
    acc = wsum(acc) + exp_b[r];
    if (lane == r) v = acc;
  }
  const float m = wmax(v);
  const float lse = logf(wsum(lane < n_roles ? expf(v - m) : 0.f));
  const float lp = v - m - lse;
  if (rows != nullptr && lane < n_roles) rows[((size_t)p * n_steps + t) * n_roles + lane] = lp;
  const uint32_t rem = remain[p];
  if (rem == 0) return;                                  // every role placed: the problem is finished (sort_model.py:151-153)
  float best = -INFINITY;
  int best_i = -1, best_role = 0;
  for (int i = 0; i < L; ++i) {
    if (!((rem >> i) & 1u)) continue;
    int role = (int)roles[(size_t)p * L + i];
    role = role < 0 ? 0 : (role >= n_roles ? n_roles - 1 : role);
    const float c = __shfl_sync(0xffffffffu, lp, role);
    if (best_i < 0 || c > best) { best = c; best_i = i; best_role = role; }
  }
  if (lane == 0) {
    remain[p] = rem & ~(1u << best_i);
    token[p] = best_role;
    pred[(size_t)p * L + t] = best_role;
    logp[(size_t)p * L + t] = best;
  }
}

// ---------------------------------------------------------------- host side
// c = act(a . w^T + bias) (+ residual): tcgen05 f16x3 when the operand twins are given, else the fp32 FFMA twin
int lin(SortCtx* c, const float* a, const F16Pair* a_b, int lda, int K, const float* w, const F16Pair* w_b, const float* bias, float* out,
        int ldc, int M, int N, const float* residual, bool relu, cudaStream_t st) {
  GemmArgs g{};
  g.nseg = 1; g.seg[0] = {a, lda, K, K, a_b};
  g.w = w; g.ldw = K; g.bias = bias; g.wb = w_b;
  g.c = out; g.ldc = ldc; g.M = M; g.N = N;
  g.cadd = residual; g.ld_cadd = ldc;
  if (c->use_tc && a_b != nullptr && w_b != nullptr && !relu) return launch_gemm_tc(g, nullptr, st);
  g.relu = relu;
  return launch_gemm_simt(g, st);
}

int ln(const float* x, const float* w, const float* b, int rows, int d, float* y, const F16Pair* y_b, cudaStream_t st) {
  k_sort_ln<<<(rows + 3) / 4, 128, 0, st>>>(x, w, b, rows, d, y, twin_out(y_b, y_b != nullptr));
  VSR_CHECK_CUDA(cudaGetLastError());
  return VSR_OK;
}

size_t weight_floats(const VsrSortDims& d) {
  const size_t D = d.d_model, F = d.d_ff;
  const size_t layer = 3 * D * D + 3 * D + D * D + D + F * D + F + D * F + D;
  return (size_t)d.n_roles * D + (size_t)d.n_verbs * D + D * D + D + 4 * D + (size_t)SORT_NPAD * D + SORT_NPAD +
         (size_t)d.n_layers * (2 * layer + (2 + 3) * 2 * D);
}

void carve_weights(SortCtx* c) {
  const VsrSortDims& d = c->d;
  const size_t D = d.d_model, F = d.d_ff;
  float* p = c->wbuf;
  auto take = [&](size_t n) { float* r = p; p += n; return r; };
  c->sr_emb = take((size_t)d.n_roles * D); c->v_emb = take((size_t)d.n_verbs * D);
  c->fc_w = take(D * D); c->fc_b = take(D);
  c->enc_ln_w = take(D); c->enc_ln_b = take(D); c->dec_ln_w = take(D); c->dec_ln_b = take(D);
  c->exp_w = take((size_t)SORT_NPAD * D); c->exp_b = take(SORT_NPAD);
  for (int s = 0; s < 2; ++s)
    for (int l = 0; l < d.n_layers; ++l) {
      SortLayer& L = s == 0 ? c->enc[l] : c->dec[l];
      L.qkv_w = take(3 * D * D); L.qkv_b = take(3 * D); L.o_w = take(D * D); L.o_b = take(D);
      L.w1_w = take(F * D); L.w1_b = take(F); L.w2_w = take(D * F); L.w2_b = take(D);
      for (int i = 0; i < (s == 0 ? 2 : 3); ++i) { L.ln_w[i] = take(D); L.ln_b[i] = take(D); }
    }
}

int ensure_ws(SortCtx* c, int P) {
  if (P <= c->cap_P) return VSR_OK;
  if (c->ws != nullptr) { VSR_CHECK_CUDA(cudaDeviceSynchronize()); cudaFree(c->ws); c->ws = nullptr; c->cap_P = 0; }
  const VsrSortDims& d = c->d;
  const size_t D = d.d_model, F = d.d_ff, L = d.max_len, nl = d.n_layers;
  const size_t rows = (size_t)P * L, KV = (size_t)P * (L + 1);
  const size_t floats = rows * D * 4 + rows * 3 * D + rows * F + (size_t)P * SORT_NPAD + nl * (rows * 2 * D + 2 * KV * D);
  VSR_CHECK_CUDA(cudaMalloc((void**)&c->ws, floats * sizeof(float) + (size_t)P * 8 + 256));
  float* p = c->ws;
  auto take = [&](size_t n) { float* r = p; p += n; return r; };
  c->x = take(rows * D); c->y = take(rows * D); c->ctx = take(rows * D); c->prior = take(rows * D);
  c->qkv = take(rows * 3 * D); c->ff = take(rows * F); c->logits = take((size_t)P * SORT_NPAD);
  for (size_t l = 0; l < nl; ++l) { c->ckv[l] = take(rows * 2 * D); c->kc[l] = take(KV * D); c->vc[l] = take(KV * D); }
  c->token = reinterpret_cast<int32_t*>(p);
  c->remain = reinterpret_cast<uint32_t*>(p) + P;
  if (c->use_tc) {
    const int rp = (int)((rows + 127) / 128 * 128);
    VSR_TRY(make_pair(&c->y_b, rp, (int)D, false)); VSR_TRY(make_pair(&c->ctx_b, rp, (int)D, false));
    VSR_TRY(make_pair(&c->ff_b, rp, (int)F, false)); VSR_TRY(make_pair(&c->prior_b, rp, (int)D, false));
  }
  c->cap_P = P;
  return VSR_OK;
}

// operand twins of the packed projection weights (after every weight load)
int split_weights(SortCtx* c, cudaStream_t st) {
  if (!c->use_tc) return VSR_OK;
  const VsrSortDims& d = c->d;
  const int D = d.d_model, F = d.d_ff;
  for (int s = 0; s < 2; ++s)
    for (int l = 0; l < d.n_layers; ++l) {
      SortLayer& L = s == 0 ? c->enc[l] : c->dec[l];
      if (L.qkv_p.hi == nullptr) {
        VSR_TRY(make_pair(&L.qkv_p, 3 * D, D, true)); VSR_TRY(make_pair(&L.o_p, D, D, true));
        VSR_TRY(make_pair(&L.w1_p, F, D, true)); VSR_TRY(make_pair(&L.w2_p, D, F, true));
        VSR_TRY(make_view(&L.q_v, L.qkv_p, 0, D)); VSR_TRY(make_view(&L.kv_v, L.qkv_p, D, 2 * D));
      }
      VSR_TRY(launch_split_pair(L.qkv_w, L.qkv_p, (size_t)3 * D * D, st, true));
      VSR_TRY(launch_split_pair(L.o_w, L.o_p, (size_t)D * D, st, true));
      VSR_TRY(launch_split_pair(L.w1_w, L.w1_p, (size_t)F * D, st, true));
      VSR_TRY(launch_split_pair(L.w2_w, L.w2_p, (size_t)D * F, st, true));
    }
  return VSR_OK;
}

// one encoder / decoder layer over `rows` rows of c->x (in place).  Decoder: one new position (step t) per problem.
int run_layer(SortCtx* c, const SortLayer& W, bool decoder, int l, int P, int t, cudaStream_t st) {
  const VsrSortDims& d = c->d;
  const int D = d.d_model, F = d.d_ff, L = d.max_len, H = d.n_heads, hd = D / H;
  const int rows = decoder ? P : P * L;
  const float isq = 1.f / sqrtf((float)hd);
  const bool tc = c->use_tc;
  const F16Pair *y_b = tc ? &c->y_b : nullptr, *ctx_b = tc ? &c->ctx_b : nullptr, *ff_b = tc ? &c->ff_b : nullptr;
  float* y32 = tc ? nullptr : c->y;             // fp32 copies of the projection inputs only for the FFMA twin
  float* ctx32 = tc ? nullptr : c->ctx;
  VSR_TRY(ln(c->x, W.ln_w[0], W.ln_b[0], rows, D, y32, y_b, st));
  VSR_TRY(lin(c, c->y, y_b, D, D, W.qkv_w, &W.qkv_p, W.qkv_b, c->qkv, 3 * D, rows, 3 * D, nullptr, false, st));
  SortAttn a{};
  a.q = c->qkv; a.ldq = 3 * D; a.out = ctx32; a.out_tw = twin_out(ctx_b, tc); a.ldo = D; a.hd = hd; a.inv_sqrt_hd = isq; a.store_pos = -1;
  if (!decoder) {
    a.nq = L; a.k = c->qkv + D; a.v = c->qkv + 2 * D; a.ldkv = 3 * D; a.kv_rows = L; a.k0 = 0; a.k1 = L;
  } else {
    a.nq = 1; a.k = c->kc[l]; a.v = c->vc[l]; a.kc = c->kc[l]; a.vc = c->vc[l]; a.ldkv = D; a.kv_rows = L + 1;
    a.store_pos = t; a.k0 = t == 0 ? 0 : 1; a.k1 = t + 1;
  }
  k_sort_attn<<<P, 32 * H, 0, st>>>(a);
  VSR_CHECK_CUDA(cudaGetLastError());
  VSR_TRY(lin(c, c->ctx, ctx_b, D, D, W.o_w, &W.o_p, W.o_b, c->x, D, rows, D, c->x, false, st));
  int nln = 1;
  if (decoder) {   // "cross" attention through the SAME attention module (sort_modules.py:88) over the encoder states
    VSR_TRY(ln(c->x, W.ln_w[1], W.ln_b[1], rows, D, y32, y_b, st));
    VSR_TRY(lin(c, c->y, y_b, D, D, W.qkv_w, &W.q_v, W.qkv_b, c->qkv, 3 * D, rows, D, nullptr, false, st));
    SortAttn x{};
    x.q = c->qkv; x.ldq = 3 * D; x.nq = 1; x.k = c->ckv[l]; x.v = c->ckv[l] + D; x.ldkv = 2 * D; x.kv_rows = L; x.k0 = 0; x.k1 = L;
    x.store_pos = -1; x.out = ctx32; x.out_tw = twin_out(ctx_b, tc); x.ldo = D; x.hd = hd; x.inv_sqrt_hd = isq;
    k_sort_attn<<<P, 32 * H, 0, st>>>(x);
    VSR_CHECK_CUDA(cudaGetLastError());
    VSR_TRY(lin(c, c->ctx, ctx_b, D, D, W.o_w, &W.o_p, W.o_b, c->x, D, rows, D, c->x, false, st));
    nln = 2;
  }
  VSR_TRY(ln(c->x, W.ln_w[nln], W.ln_b[nln], rows, D, y32, y_b, st));
  if (tc) {       // the ReLU sits between two tensor-core projections: its kernel writes the second one's operand twins
    VSR_TRY(lin(c, c->y, y_b, D, D, W.w1_w, &W.w1_p, W.w1_b, c->ff, F, rows, F, nullptr, false, st));
    const size_t n4 = (size_t)rows * F / 4;
    k_relu_twin<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(c->ff, n4, twin_out(ff_b));
    VSR_CHECK_CUDA(cudaGetLastError());
  } else {
    VSR_TRY(lin(c, c->y, nullptr, D, D, W.w1_w, nullptr, W.w1_b, c->ff, F, rows, F, nullptr, true, st));
  }
  VSR_TRY(lin(c, c->ff, ff_b, F, F, W.w2_w, &W.w2_p, W.w2_b, c->x, D, rows, D, c->x, false, st));
  return VSR_OK;
}

int generate_impl(SortCtx* c, const int64_t* verbs, const int64_t* roles, int P, int n_steps, const int32_t* n_active, int64_t* pred,
                  float* logp, float* step_rows, cudaStream_t st) {
  const VsrSortDims& d = c->d;
  const int D = d.d_model, L = d.max_len;
  const float scale = sqrtf((float)D);
  VSR_TRY(ensure_ws(c, P));
  VSR_CHECK_CUDA(cudaMemsetAsync(pred, 0, sizeof(int64_t) * (size_t)P * L, st));
  VSR_CHECK_CUDA(cudaMemsetAsync(logp, 0, sizeof(float) * (size_t)P * L, st));
  k_sort_init<<<(P + 127) / 128, 128, 0, st>>>(roles, P, L, c->remain, c->token);
  VSR_CHECK_CUDA(cudaGetLastError());
  // ---- encoder over the L role positions of every problem
  k_sort_embed_enc<<<P * L, 128, 0, st>>>(verbs, roles, c->v_emb, c->sr_emb, L, D, d.n_verbs, d.n_roles, scale, c->x);
  VSR_CHECK_CUDA(cudaGetLastError());
  if (d.add_fc) {
    VSR_TRY(lin(c, c->x, nullptr, D, D, c->fc_w, nullptr, c->fc_b, c->y, D, P * L, D, nullptr, false, st));
    VSR_CHECK_CUDA(cudaMemcpyAsync(c->x, c->y, sizeof(float) * (size_t)P * L * D, cudaMemcpyDeviceToDevice, st));
  }
  for (int l = 0; l < d.n_layers; ++l) VSR_TRY(run_layer(c, c->enc[l], false, l, P, 0, st));
  VSR_TRY(ln(c->x, c->enc_ln_w, c->enc_ln_b, P * L, D, c->use_tc ? nullptr : c->prior, c->use_tc ? &c->prior_b : nullptr, st));
  // keys / values of the encoder states under each decoder layer's attention module
  for (int l = 0; l < d.n_layers; ++l)
    VSR_TRY(lin(c, c->prior, c->use_tc ? &c->prior_b : nullptr, D, D, c->dec[l].qkv_w + (size_t)D * D, &c->dec[l].kv_v, c->dec[l].qkv_b + D,
                c->ckv[l], 2 * D, P * L, 2 * D, nullptr, false, st));
  // ---- decoder: one position per step, greedy over the roles still to be placed
  for (int t = 0; t < n_steps; ++t) {
    // problems sorted by falling role count: only the first n_active[t] still have a role to place at step t
    const int Pt = n_active != nullptr ? n_active[t] : P;
    if (Pt <= 0) break;
    k_sort_embed_dec<<<Pt, 128, 0, st>>>(c->token, c->sr_emb, D, scale, c->x);
    VSR_CHECK_CUDA(cudaGetLastError());
    for (int l = 0; l < d.n_layers; ++l) VSR_TRY(run_layer(c, c->dec[l], true, l, Pt, t, st));
    k_sort_select<<<(Pt + 7) / 8, 256, sizeof(float) * d.n_roles * D, st>>>(c->x, D, c->dec_ln_w, c->dec_ln_b, c->exp_w, c->exp_b, d.n_roles,
                                                                            roles, L, Pt, t, n_steps, c->remain, c->token, pred, logp,
                                                                            step_rows);
    VSR_CHECK_CUDA(cudaGetLastError());
  }
  return VSR_OK;
}

int n_weight_ptrs(const VsrSortDims& d) { return 10 + 34 * d.n_layers; }

}  // namespace
}  // namespace vsr

using vsr::SortCtx;

extern "C" {

int vsr_sort_load_weights(vsr_sort_handle h, const float* const* weights, int32_t n_weights, void* stream) {
  if (!h || !weights) { vsr::set_error("vsr_sort_load_weights: null argument"); return VSR_EINVAL; }
  SortCtx* c = (SortCtx*)h;
  const VsrSortDims& d = c->d;
  VSR_REQUIRE(n_weights == vsr::n_weight_ptrs(d), VSR_EINVAL, "vsr_sort_load_weights: expected %d weight pointers, got %d",
              vsr::n_weight_ptrs(d), n_weights);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t D = d.d_model, F = d.d_ff;
  int i = 0;
  int rc = VSR_OK;
  auto cp = [&](float* dst, size_t n) {
    const float* src = weights[i++];
    if (rc != VSR_OK) return;
    if (src == nullptr) { vsr::set_error("vsr_sort_load_weights: weight %d is null", i - 1); rc = VSR_EINVAL; return; }
    if (cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
      vsr::set_error("vsr_sort_load_weights: copy of weight %d failed", i - 1); rc = VSR_ECUDA;
    }
  };
  cp(c->sr_emb, (size_t)d.n_roles * D); cp(c->v_emb, (size_t)d.n_verbs * D);
  if (d.add_fc) { cp(c->fc_w, D * D); cp(c->fc_b, D); } else i += 2;
  cp(c->enc_ln_w, D); cp(c->enc_ln_b, D);
  for (int s = 0; s < 2; ++s) {
    if (s == 1) { cp(c->dec_ln_w, D); cp(c->dec_ln_b, D); }
    for (int l = 0; l < d.n_layers; ++l) {
      vsr::SortLayer& L = s == 0 ? c->enc[l] : c->dec[l];
      for (int m = 0; m < 3; ++m) { cp(L.qkv_w + m * D * D, D * D); cp(L.qkv_b + m * D, D); }    // linear_Q, linear_K, linear_V
      cp(L.o_w, D * D); cp(L.o_b, D);
      cp(L.w1_w, F * D); cp(L.w1_b, F); cp(L.w2_w, D * F); cp(L.w2_b, D);
      for (int n = 0; n < (s == 0 ? 2 : 3); ++n) { cp(L.ln_w[n], D); cp(L.ln_b[n], D); }
    }
  }
  cp(c->exp_w, (size_t)d.n_roles * D); cp(c->exp_b, d.n_roles);      // rows n_roles..63 stay zero
  if (rc != VSR_OK) return rc;
  return vsr::split_weights(c, st);
}

int vsr_sort_create(const VsrSortDims* dims, const float* const* weights, int32_t n_weights, vsr_sort_handle* out) {
  if (!out || !weights || !dims) { vsr::set_error("vsr_sort_create: null argument"); return VSR_EINVAL; }
  *out = nullptr;
  const VsrSortDims& d = *dims;
  VSR_REQUIRE(d.n_layers >= 1 && d.n_layers <= vsr::SORT_MAXL, VSR_EINVAL, "vsr_sort_create: n_layers=%d not in [1,%d]", d.n_layers, vsr::SORT_MAXL);
  VSR_REQUIRE(d.max_len >= 1 && d.max_len + 1 <= vsr::SORT_MAXK, VSR_EINVAL, "vsr_sort_create: max_len=%d not in [1,%d]", d.max_len, vsr::SORT_MAXK - 1);
  VSR_REQUIRE(d.n_roles >= 1 && d.n_roles <= 32 && d.n_verbs >= 1, VSR_EINVAL, "vsr_sort_create: n_roles=%d must be <= 32", d.n_roles);
  VSR_REQUIRE(d.n_heads >= 1 && d.n_heads <= 32 && d.d_model % d.n_heads == 0 && d.d_model % 64 == 0 && d.d_ff % 64 == 0, VSR_EINVAL,
              "vsr_sort_create: d_model=%d / d_ff=%d must be multiples of 64 and d_model of n_heads=%d", d.d_model, d.d_ff, d.n_heads);
  VSR_REQUIRE(d.d_model <= 32 * vsr::SORT_DPL, VSR_EINVAL, "vsr_sort_create: d_model=%d > %d", d.d_model, 32 * vsr::SORT_DPL);
  VSR_REQUIRE((d.d_model / d.n_heads) % 4 == 0 && d.d_model / d.n_heads <= 4 * vsr::SORT_HD4, VSR_EINVAL,
              "vsr_sort_create: head dimension %d must be a multiple of 4 and at most %d", d.d_model / d.n_heads, 4 * vsr::SORT_HD4);
  int ndev = 0;
  VSR_REQUIRE(cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0, VSR_ECUDA, "vsr_sort_create: no CUDA device (no CPU fallback)");
  SortCtx* c = new SortCtx();
  c->d = d;
  if (const char* e = getenv("VSRDEC_GEMM")) c->use_tc = strcmp(e, "simt") != 0;
  VSR_CHECK_CUDA(cudaGetDevice(&c->device));
  VSR_REQUIRE(sizeof(float) * d.n_roles * d.d_model <= 200 * 1024, VSR_EINVAL, "vsr_sort_create: role projection does not fit shared memory");
  VSR_CHECK_CUDA(cudaFuncSetAttribute(vsr::k_sort_select, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const size_t n = vsr::weight_floats(d);
  if (cudaMalloc((void**)&c->wbuf, n * sizeof(float)) != cudaSuccess) { vsr::set_error("vsr_sort_create: cudaMalloc failed"); delete c; return VSR_ENOMEM; }
  cudaMemset(c->wbuf, 0, n * sizeof(float));
  vsr::carve_weights(c);
  const int r = vsr_sort_load_weights((vsr_sort_handle)c, weights, n_weights, nullptr);
  if (r != VSR_OK) { cudaFree(c->wbuf); delete c; return r; }
  VSR_CHECK_CUDA(cudaStreamSynchronize(nullptr));
  *out = (vsr_sort_handle)c;
  return VSR_OK;
}

void vsr_sort_destroy(vsr_sort_handle h) {
  if (!h) return;
  SortCtx* c = (SortCtx*)h;
  cudaDeviceSynchronize();
  cudaFree(c->wbuf);
  if (c->ws) cudaFree(c->ws);
  for (int s = 0; s < 2; ++s)
    for (int l = 0; l < vsr::SORT_MAXL; ++l) {
      vsr::SortLayer& L = s == 0 ? c->enc[l] : c->dec[l];
      vsr::free_pair(&L.q_v, true); vsr::free_pair(&L.kv_v, true);
      vsr::free_pair(&L.qkv_p); vsr::free_pair(&L.o_p); vsr::free_pair(&L.w1_p); vsr::free_pair(&L.w2_p);
    }
  vsr::free_pair(&c->y_b); vsr::free_pair(&c->ctx_b); vsr::free_pair(&c->ff_b); vsr::free_pair(&c->prior_b);
  delete c;
}

int vsr_sort_generate(vsr_sort_handle h, const int64_t* verbs, const int64_t* roles, int32_t P, int32_t n_steps,
                      const int32_t* n_active, int64_t* pred, float* logp, float* step_rows, void* stream) {
  if (!h) { vsr::set_error("vsr_sort_generate: null handle"); return VSR_EINVAL; }
  SortCtx* c = (SortCtx*)h;
  VSR_REQUIRE(verbs && roles && pred && logp && P >= 0, VSR_EINVAL, "vsr_sort_generate: null argument");
  VSR_REQUIRE(n_steps >= 0 && n_steps <= c->d.max_len, VSR_EINVAL, "vsr_sort_generate: n_steps=%d not in [0,%d]", n_steps, c->d.max_len);
  if (P == 0) return VSR_OK;
  if (n_active != nullptr)
    for (int t = 0; t < n_steps; ++t)
      VSR_REQUIRE(n_active[t] >= 0 && n_active[t] <= P && (t == 0 || n_active[t] <= n_active[t - 1]), VSR_EINVAL,
                  "vsr_sort_generate: n_active must be non-increasing and within [0, P] (problems sorted by falling role count)");
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != c->device) cudaSetDevice(c->device);
  const int rc = vsr::generate_impl(c, verbs, roles, P, n_steps, n_active, pred, logp, step_rows, (cudaStream_t)stream);
  if (prev != c->device && prev >= 0) cudaSetDevice(prev);
  return rc;
}

}  // extern "C"
