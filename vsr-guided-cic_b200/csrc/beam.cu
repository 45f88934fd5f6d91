// Beam expansion without host synchronisation: joint (word, gate) top-k per caption over
// cur_beam x (k word candidates) x (2 gates), sticky EOS masks, back-pointer history, beam-state
// reorder, and the final sort + back-track.
// Reference: models/CaptioningModel.py:116-195 (beam_search), :197-294 (beam_search_v),
// :78-114 (_select_beam).  Statics are never copied per beam: rows index
// them by caption = row / beam.
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"

namespace vsr {

namespace {

__device__ __forceinline__ bool before(float v1, int i1, float v2, int i2) {
  return v1 > v2 || (v1 == v2 && i1 < i2);
}

// log-prob of vocabulary entry `word` in row `row` as returned by the step (post verb forcing)
__device__ __forceinline__ float row_logp(const float* logits, int ld, const float* row_max,
                                          const float* row_lsum, const int32_t* forced, int row, int word) {
  const int f = forced[row];
  if (f >= 0) return word == f ? 0.f : -1e6f;
  return (logits[(size_t)row * ld + word] - row_max[row]) - row_lsum[row];
}

struct BeamArgs {
  int t, b, cur, k, V;
  int64_t eos0, eos1;
  const float* logits; int ld;
  const float* row_max; const float* row_lsum; const int32_t* forced; const int32_t* cand;
  const float* gate_lp;
  const float* seq_lp; float* seq_lp_n;
  float *m0, *m1, *m0n, *m1n;
  const int32_t *prev_word, *prev_gate;            // previous step's picks per current slot (t > 0)
  int32_t *sel_beam, *sel_word, *sel_gate;        // this step's picks (double-buffered against prev_*)
  int32_t *hist_parent, *hist_word, *hist_gate;    // [T][b][k] slices for step t
  float *hist_score, *hist_lpw, *hist_lpg;
  const int32_t *f_beam, *f_word, *f_gate;        // forced selections for step t or null
};

constexpr int MAXC = 2 * VSR_MAX_BEAM * VSR_MAX_BEAM;  // 128 candidates per caption

struct BeamSmem {
  float seq[VSR_MAX_BEAM], m0[VSR_MAX_BEAM], m1[VSR_MAX_BEAM], ps[VSR_MAX_BEAM];
  int full[VSR_MAX_BEAM], pb[VSR_MAX_BEAM], pw[VSR_MAX_BEAM], pg[VSR_MAX_BEAM];
};

// selection for caption c by one warp; leaves the picks (parent, word, gate) in sh.pb / pw / pg
__device__ __forceinline__ void beam_select_warp(const BeamArgs& a, BeamSmem& sh, int c, int lane, bool write) {
  const int k = a.k, cur = a.cur;
  float* s_seq = sh.seq; float* s_m0 = sh.m0; float* s_m1 = sh.m1; float* s_ps = sh.ps;
  int* s_full = sh.full; int* s_pb = sh.pb; int* s_pw = sh.pw; int* s_pg = sh.pg;

  if (lane < cur) {
    float m0 = 1.f, m1 = 1.f, seq = 0.f;
    if (a.t > 0) {
      // sticky masks: a head's mask drops to 0 once the slot's previous token was its EOS (:143-144)
      seq = a.seq_lp[c * k + lane];
      m0 = a.m0[c * k + lane] * ((int64_t)a.prev_word[c * k + lane] != a.eos0 ? 1.f : 0.f);
      m1 = a.m1[c * k + lane] * ((int64_t)a.prev_gate[c * k + lane] != a.eos1 ? 1.f : 0.f);
    }
    s_seq[lane] = seq; s_m0[lane] = m0; s_m1[lane] = m1;
    s_full[lane] = (a.t == 0) ? 1 : (fminf(fmaxf(m0 + m1, 0.f), 1.f) != 0.f);
  }
  __syncwarp();

  if (a.f_beam == nullptr) {
    // candidates: idx = (parent j, i-th word candidate, gate g); lane owns idx = lane + 32*q
    float cs[MAXC / 32]; int cf[MAXC / 32]; bool used[MAXC / 32];
    const int ncand = cur * k * 2;
#pragma unroll
    for (int q = 0; q < MAXC / 32; ++q) {
      const int idx = lane + 32 * q;
      cs[q] = -INFINITY; cf[q] = 0x7fffffff; used[q] = true;
      if (idx < ncand) {
        const int j = idx / (2 * k), i = (idx >> 1) % k, g = idx & 1;
        const int row = c * cur + j;
        int word; float sc;
        if (s_full[j]) {
          word = a.cand[row * VSR_MAX_BEAM + i];
          const float wl = row_logp(a.logits, a.ld, a.row_max, a.row_lsum, a.forced, row, word);
          sc = __fadd_rn(s_seq[j], __fadd_rn(wl, a.gate_lp[row * 2 + g]));   // seq + (word + gate), :139
        } else {
          // frozen beam: old score at word 0 (both gates), -999 elsewhere (:146-150)
          word = i;
          sc = (i == 0) ? s_seq[j] : -999.f;
        }
        cs[q] = sc; cf[q] = j * 2 * a.V + word * 2 + g; used[q] = false;
      }
    }
    for (int sel = 0; sel < k; ++sel) {
      float bv = -INFINITY; int bf = 0x7fffffff;
#pragma unroll
      for (int q = 0; q < MAXC / 32; ++q)
        if (!used[q] && before(cs[q], cf[q], bv, bf)) { bv = cs[q]; bf = cf[q]; }
      {   // warp arg-best in (score desc, flat index asc) order: two redux.sync + one broadcast
        const unsigned u = __float_as_uint(bv);
        const unsigned key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        const unsigned kbest = __reduce_max_sync(0xffffffffu, key);
        const int fbest = (int)__reduce_min_sync(0xffffffffu, key == kbest ? (unsigned)bf : 0xffffffffu);
        const unsigned owner = __ballot_sync(0xffffffffu, key == kbest && bf == fbest);
        bv = __shfl_sync(0xffffffffu, bv, __ffs(owner) - 1);
        bf = fbest;
      }
#pragma unroll
      for (int q = 0; q < MAXC / 32; ++q) if (cf[q] == bf) used[q] = true;
      if (lane == 0) {
        const int j = bf / (2 * a.V), rem = bf - j * 2 * a.V;
        s_pb[sel] = j; s_pw[sel] = rem >> 1; s_pg[sel] = rem & 1; s_ps[sel] = bv;
      }
    }
  } else if (lane < k) {
    // trajectory replay: take the given selection, score it with this run's own numbers
    const int j = a.f_beam[c * k + lane], word = a.f_word[c * k + lane], g = a.f_gate[c * k + lane];
    const int row = c * cur + j;
    float sc;
    if (s_full[j]) {
      const float wl = row_logp(a.logits, a.ld, a.row_max, a.row_lsum, a.forced, row, word);
      sc = __fadd_rn(s_seq[j], __fadd_rn(wl, a.gate_lp[row * 2 + g]));
    } else {
      sc = (word == 0) ? s_seq[j] : -999.f;
    }
    s_pb[lane] = j; s_pw[lane] = word; s_pg[lane] = g; s_ps[lane] = sc;
  }
  __syncwarp();

  if (lane < k && write) {
    const int j = s_pb[lane], word = s_pw[lane], g = s_pg[lane];
    const int row = c * cur + j;
    const int o = c * k + lane;
    // per-token log-probs of the pick, in slot order, masked by the parent's sticky masks (:145,175-177)
    float lw = row_logp(a.logits, a.ld, a.row_max, a.row_lsum, a.forced, row, word);
    float lg = a.gate_lp[row * 2 + g];
    if (a.t > 0) { lw *= s_m0[j]; lg *= s_m1[j]; }
    a.sel_beam[o] = j; a.sel_word[o] = word; a.sel_gate[o] = g;
    a.seq_lp_n[o] = s_ps[lane];
    a.m0n[o] = s_m0[j]; a.m1n[o] = s_m1[j];
    a.hist_parent[o] = j; a.hist_word[o] = word; a.hist_gate[o] = g;
    a.hist_score[o] = s_ps[lane]; a.hist_lpw[o] = lw; a.hist_lpg[o] = lg;
  }
}

// ---------------------------------------------------------------- state advance / reorder
struct AdvanceArgs {
  int rows_new, cur, k;            // new row n' = c*k + i ; parent row = c*cur + parent[n'] (or n' if null)
  const int32_t* parent;
  const int32_t* word32;           // next input token per new row ...
  const int64_t* word64; int64_t word64_stride;   // ... or int64 strided (teacher forcing)
  const int32_t* gate32;           // slot shift per new row, or null
  int fixed_slot;                  // >= 0: set the pointer to this slot (teacher forcing)
  int L, Hp, Ep, V;
  const float *h1n, *c1n, *h2n, *c2n;
  float *h1, *c1, *h2, *c2;
  const int32_t* ptr; int32_t* ptrn;
  int32_t* word_idx;
  // tensor-core twins of h1 / h2 (none on the fp32 path), moved as raw 16-byte vectors: per state up to three arrays
  // (fp16 hi + fp16 lo, or fp16 hi + e4m3 hi8 + e4m3 lo8) of `tw_vec[i]` vectors per row
  int n_tw;
  const uint4* tw_src[8]; uint4* tw_dst[8]; int tw_vec[8];
};

// copy the state of parent row p into new row n, record the next input word, update the slot pointer
__device__ __forceinline__ void advance_row(const AdvanceArgs& a, int n, int p, int64_t w, int gate_shift) {
  const size_t so = (size_t)p * a.Hp, dof = (size_t)n * a.Hp;
  for (int i = threadIdx.x * 4; i < a.Hp; i += blockDim.x * 4) {
    if (a.h1n != nullptr) {        // fp32 h only where somebody reads it (FFMA twin)
      *reinterpret_cast<float4*>(a.h1 + dof + i) = *reinterpret_cast<const float4*>(a.h1n + so + i);
      *reinterpret_cast<float4*>(a.h2 + dof + i) = *reinterpret_cast<const float4*>(a.h2n + so + i);
    }
    *reinterpret_cast<float4*>(a.c1 + dof + i) = *reinterpret_cast<const float4*>(a.c1n + so + i);
    *reinterpret_cast<float4*>(a.c2 + dof + i) = *reinterpret_cast<const float4*>(a.c2n + so + i);
  }
  w = w < 0 ? 0 : (w >= a.V ? a.V - 1 : w);
  for (int q = 0; q < a.n_tw; ++q) {
    const int nv = a.tw_vec[q];
    const uint4* src = a.tw_src[q] + (size_t)p * nv;
    uint4* dst = a.tw_dst[q] + (size_t)n * nv;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) dst[i] = src[i];
  }
  if (threadIdx.x == 0) {
    a.word_idx[n] = (int32_t)w;
    int s;
    if (a.fixed_slot >= 0) s = a.fixed_slot;
    else {
      s = a.ptr[p] + gate_shift;                               // ctrl_det_idxs + prev gate, clamped
      s = s < 0 ? 0 : (s > a.L - 1 ? a.L - 1 : s);             // (controllable_captioning.py:139-140)
    }
    a.ptrn[n] = s;
  }
}

__global__ void __launch_bounds__(256) k_advance(const AdvanceArgs a) {
  const int n = blockIdx.x;
  const int c = n / a.k;
  const int p = a.parent != nullptr ? c * a.cur + a.parent[n] : n;
  const int64_t w = a.word32 != nullptr ? (int64_t)a.word32[n] : a.word64[(size_t)n * a.word64_stride];
  advance_row(a, n, p, w, a.gate32 != nullptr ? a.gate32[n] : 0);
}

// beam selection fused with the reorder of the beam states.  Grid (captions, k): every CTA of a caption
// repeats the (cheap, deterministic) selection with its warp 0 — only CTA y = 0 records it — and then
// moves ONE new beam row, so the copy keeps captions x k CTAs of parallelism without a second launch.
__global__ void __launch_bounds__(256) k_beam_step(const BeamArgs b, const AdvanceArgs a, int do_advance) {
  __shared__ BeamSmem sh;
  const int c = blockIdx.x, i = blockIdx.y;
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x < 32) beam_select_warp(b, sh, c, threadIdx.x, i == 0);
  __syncthreads();
  if (!do_advance) return;
  advance_row(a, c * b.k + i, c * b.cur + sh.pb[i], (int64_t)sh.pw[i], sh.pg[i]);
}

// zero state, slot 0, input word = bos   (init_state, controllable_captioning.py:109-115, :136)
__global__ void k_state_init(float* h1, float* c1, float* h2, float* c2, int32_t* word_idx, int32_t* ptr,
                             int bos, int Hp, TwinOut h1b, TwinOut h2b) {
  const int n = blockIdx.x;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = threadIdx.x * 4; i < Hp; i += blockDim.x * 4) {
    const size_t o = (size_t)n * Hp + i;
    *reinterpret_cast<float4*>(h1 + o) = z; *reinterpret_cast<float4*>(c1 + o) = z;
    *reinterpret_cast<float4*>(h2 + o) = z; *reinterpret_cast<float4*>(c2 + o) = z;
    store_twin4(h1b, o, z); store_twin4(h2b, o, z);       // (all-zero bit patterns in fp16 and e4m3 alike)
  }
  if (threadIdx.x == 0) { ptr[n] = 0; word_idx[n] = bos; }
}

__global__ void k_words_from_i64(const int64_t* words, int32_t* word_idx, int rows, int V) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= rows) return;
  int64_t w = words[n];
  w = w < 0 ? 0 : (w >= V ? V - 1 : w);
  word_idx[n] = (int32_t)w;
}

// greedy pick of both heads (CaptioningModel.test, CaptioningModel.py:47): first maximum wins
__global__ void k_greedy_pick(const int32_t* cand, const float* gate_lp, int32_t* sel_word,
                              int32_t* sel_gate, int64_t* out_words, int64_t* out_gates, int rows,
                              int t, int T) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= rows) return;
  const int w = cand[n * VSR_MAX_BEAM];
  const int g = gate_lp[n * 2 + 1] > gate_lp[n * 2 + 0] ? 1 : 0;
  sel_word[n] = w; sel_gate[n] = g;
  out_words[(size_t)n * T + t] = w; out_gates[(size_t)n * T + t] = g;
}

// ---------------------------------------------------------------- multinomial sampling (CaptioningModel.sample_rl)
// Philox4x32-10 counter-based generator: the draw of (seed, row, step, vocabulary entry) never depends on launch shape.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ float u01(unsigned x) { return ((float)(x >> 8) + 0.5f) * (1.f / 16777216.f); }   // (0, 1)

// One CTA per row: word ~ Categorical(softmax(logits)) by the Gumbel-max trick (argmax of logit + G, G = -log(-log u),
// which draws index v with probability exp(logit_v) / sum exp), gate ~ Categorical(exp(gate_lp)); log-probs of the
// draws as torch.distributions.Categorical(logits=out).log_prob(sample) returns them (CaptioningModel.py:66-70).
__global__ void __launch_bounds__(256) k_sample_pick(const float* __restrict__ logits, int ld, int V, const float* __restrict__ row_max,
                                                     const float* __restrict__ row_lsum, const float* __restrict__ gate_lp,
                                                     unsigned long long seed, int t, int T, int32_t* sel_word, int32_t* sel_gate,
                                                     int64_t* out_words, int64_t* out_gates, float* lp_words, float* lp_gates) {
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* x = logits + (size_t)n * ld;
  const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
  float bv = -INFINITY; int bi = 0x7fffffff;
  for (int v0 = tid * 4; v0 < V; v0 += 256 * 4) {
    const uint4 r = philox4x32_10(make_uint4((unsigned)(v0 >> 2), (unsigned)n, (unsigned)t, 0u), key);
    const float4 q = *reinterpret_cast<const float4*>(x + v0);        // rows are padded to a multiple of 4 columns
    const float xs[4] = {q.x, q.y, q.z, q.w};
    const unsigned rs[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (v0 + e >= V) break;
      const float g = -logf(-logf(u01(rs[e])));
      const float sc = xs[e] + g;
      if (sc > bv) { bv = sc; bi = v0 + e; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (before(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) { s_v[warp] = bv; s_i[warp] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) if (before(s_v[w], s_i[w], bv, bi)) { bv = s_v[w]; bi = s_i[w]; }
    const int word = bi < V ? bi : V - 1;
    const float lw = (x[word] - row_max[n]) - row_lsum[n];
    // gate: two categories with log-probs gate_lp[n]; counter word 1 keeps this draw apart from the vocabulary's
    const uint4 r = philox4x32_10(make_uint4(0u, (unsigned)n, (unsigned)t, 1u), key);
    const float g0 = gate_lp[(size_t)n * 2], g1 = gate_lp[(size_t)n * 2 + 1];
    const int gate = u01(r.x) < expf(g0) ? 0 : 1;
    sel_word[n] = word; sel_gate[n] = gate;
    out_words[(size_t)n * T + t] = word; out_gates[(size_t)n * T + t] = gate;
    lp_words[(size_t)n * T + t] = lw; lp_gates[(size_t)n * T + t] = gate == 0 ? g0 : g1;
  }
}

// final ordering of the beams by accumulated score (stable, descending) and unroll of the
// back-pointers (CaptioningModel.py:182-194 / 279-293).  One CTA per caption: the caption's [T][k] history is
// staged in shared memory with coalesced loads, so the serial pointer chase never waits on global memory.
__global__ void __launch_bounds__(64) k_backtrack(int b, int k, int T, int out_size, const float* seq_lp,
                                                  const int32_t* hist_parent, const int32_t* hist_word,
                                                  const int32_t* hist_gate, const float* hist_lpw, const float* hist_lpg,
                                                  int64_t* out_words, int64_t* out_gates, float* lp_words, float* lp_gates) {
  extern __shared__ __align__(16) unsigned char bt_smem[];      // 5 arrays of T*k 4-byte entries (<= 40 KB)
  int32_t* s_par = reinterpret_cast<int32_t*>(bt_smem);
  int32_t* s_wrd = s_par + T * k;
  int32_t* s_gat = s_wrd + T * k;
  float* s_lpw = reinterpret_cast<float*>(s_gat + T * k);
  float* s_lpg = s_lpw + T * k;
  __shared__ float s_seq[VSR_MAX_BEAM];
  const int c = blockIdx.x;
  for (int i = threadIdx.x; i < T * k; i += blockDim.x) {
    const int t = i / k, j = i - t * k;
    const size_t hi = ((size_t)t * b + c) * k + j;
    s_par[i] = hist_parent[hi]; s_wrd[i] = hist_word[hi]; s_gat[i] = hist_gate[hi];
    s_lpw[i] = hist_lpw[hi]; s_lpg[i] = hist_lpg[hi];
  }
  if (threadIdx.x < k) s_seq[threadIdx.x] = seq_lp[c * k + threadIdx.x];
  __syncthreads();
  const int o = threadIdx.x;
  if (o >= out_size) return;
  int order[VSR_MAX_BEAM];
  for (int i = 0; i < k; ++i) order[i] = i;
  for (int i = 1; i < k; ++i) {   // insertion sort, stable, descending
    const int oi = order[i];
    const float v = s_seq[oi];
    int j = i - 1;
    while (j >= 0 && s_seq[order[j]] < v) { order[j + 1] = order[j]; --j; }
    order[j + 1] = oi;
  }
  const int final_slot = order[o];
  int slot = final_slot;
  const size_t ob = ((size_t)c * out_size + o) * T;
  for (int t = T - 1; t >= 0; --t) {
    out_words[ob + t] = s_wrd[t * k + slot];
    out_gates[ob + t] = s_gat[t * k + slot];
    // log-probs are NOT back-tracked: slot order at step t, permuted by the final sort only
    lp_words[ob + t] = s_lpw[t * k + final_slot];
    lp_gates[ob + t] = s_lpg[t * k + final_slot];
    slot = s_par[t * k + slot];
  }
}

}  // namespace

static TwinOut pp(const Ctx* c, const F16Pair& b) { return twin_out(&b, c->use_tc); }

int launch_state_init(Ctx* c, int rows, cudaStream_t st) {
  k_state_init<<<rows, 256, 0, st>>>(c->h1, c->c1, c->h2, c->c2, c->word_idx, c->ptr, c->d.bos_idx, c->Hp,
                                     pp(c, c->h1_b), pp(c, c->h2_b));
  VSR_CHECK_CUDA(cudaGetLastError()); c->launches++;
  return VSR_OK;
}

int launch_words(Ctx* c, const int64_t* words, int rows, cudaStream_t st) {
  k_words_from_i64<<<(rows + 127) / 128, 128, 0, st>>>(words, c->word_idx, rows, c->V);
  VSR_CHECK_CUDA(cudaGetLastError()); c->launches++;
  return VSR_OK;
}

static void fill_advance(Ctx* c, AdvanceArgs& a) {
  a.L = c->L; a.Hp = c->Hp; a.Ep = c->Ep; a.V = c->V;
  a.h1n = c->state_h32 ? c->h1n : nullptr; a.c1n = c->c1n; a.h2n = c->state_h32 ? c->h2n : nullptr; a.c2n = c->c2n;
  a.h1 = c->h1; a.c1 = c->c1; a.h2 = c->h2; a.c2 = c->c2;
  a.ptr = c->ptr; a.ptrn = c->ptrn; a.word_idx = c->word_idx;
  a.n_tw = 0;
  if (c->use_tc) {
    auto add = [&](const void* src, void* dst, int bytes_per_elem) {
      if (src == nullptr || dst == nullptr) return;
      a.tw_src[a.n_tw] = (const uint4*)src; a.tw_dst[a.n_tw] = (uint4*)dst; a.tw_vec[a.n_tw] = c->Hp * bytes_per_elem / 16;
      ++a.n_tw;
    };
    const F16Pair* src[2] = {&c->h1n_b, &c->h2n_b};
    F16Pair* dst[2] = {&c->h1_b, &c->h2_b};
    for (int q = 0; q < 2; ++q) {
      add(src[q]->hi, dst[q]->hi, 2); add(src[q]->lo, dst[q]->lo, 2);
      add(src[q]->hi8, dst[q]->hi8, 1); add(src[q]->lo8, dst[q]->lo8, 1);
    }
  }
}

// beam selection of step t and (unless it is the last step) the state reorder for step t+1, one launch
int launch_beam_step(Ctx* c, int t, int b, int cur, int k, int64_t eos0, int64_t eos1, const int32_t* f_beam,
                     const int32_t* f_word, const int32_t* f_gate, bool advance, cudaStream_t st) {
  PhaseScope ps(c, PH_BEAM, st);
  BeamArgs a{};
  a.t = t; a.b = b; a.cur = cur; a.k = k; a.V = c->V; a.eos0 = eos0; a.eos1 = eos1;
  a.logits = c->logits; a.ld = c->NE; a.row_max = c->row_max; a.row_lsum = c->row_lsum;
  a.forced = c->forced; a.cand = c->cand; a.gate_lp = c->gate_lp;
  a.seq_lp = c->seq_lp; a.seq_lp_n = c->seq_lp_n;
  a.m0 = c->m0; a.m1 = c->m1; a.m0n = c->m0n; a.m1n = c->m1n;
  a.prev_word = c->sel_word; a.prev_gate = c->sel_gate;
  a.sel_beam = c->sel_beam_n; a.sel_word = c->sel_word_n; a.sel_gate = c->sel_gate_n;
  const size_t off = (size_t)t * b * k;
  a.hist_parent = c->hist_parent + off; a.hist_word = c->hist_word + off; a.hist_gate = c->hist_gate + off;
  a.hist_score = c->hist_score + off; a.hist_lpw = c->hist_lpw + off; a.hist_lpg = c->hist_lpg + off;
  a.f_beam = f_beam; a.f_word = f_word; a.f_gate = f_gate;
  AdvanceArgs ad{};
  ad.rows_new = b * k; ad.cur = cur; ad.k = k; ad.fixed_slot = -1;
  fill_advance(c, ad);
  VSR_CHECK_CUDA(launch_k(k_beam_step, dim3(b, advance ? k : 1), dim3(256), 0, st, c->use_pdl && (c->pdl_mode & 2), a, ad, advance ? 1 : 0));
  VSR_CHECK_CUDA(cudaGetLastError()); c->launches++;
  std::swap(c->sel_beam, c->sel_beam_n);
  std::swap(c->sel_word, c->sel_word_n);
  std::swap(c->sel_gate, c->sel_gate_n);
  std::swap(c->seq_lp, c->seq_lp_n);
  std::swap(c->m0, c->m0n);
  std::swap(c->m1, c->m1n);
  if (advance) std::swap(c->ptr, c->ptrn);
  return VSR_OK;
}

static int launch_advance(Ctx* c, AdvanceArgs& a, cudaStream_t st) {
  fill_advance(c, a);
  k_advance<<<a.rows_new, 256, 0, st>>>(a);
  VSR_CHECK_CUDA(cudaGetLastError()); c->launches++;
  std::swap(c->ptr, c->ptrn);
  return VSR_OK;
}

int launch_commit_identity(Ctx* c, int rows, const int64_t* next_words, int64_t word_stride,
                           int next_slot, cudaStream_t st) {
  PhaseScope ps(c, PH_REORDER, st);
  AdvanceArgs a{};
  a.rows_new = rows; a.cur = 1; a.k = 1;
  a.parent = nullptr;
  if (next_words != nullptr) { a.word64 = next_words; a.word64_stride = word_stride; a.gate32 = nullptr; }
  else { a.word32 = c->sel_word; a.gate32 = c->sel_gate; }
  a.fixed_slot = next_slot;
  return launch_advance(c, a, st);
}

int launch_greedy_pick(Ctx* c, int rows, int t, int T, int64_t* out_words, int64_t* out_gates,
                       cudaStream_t st) {
  PhaseScope ps(c, PH_BEAM, st);
  k_greedy_pick<<<(rows + 127) / 128, 128, 0, st>>>(c->cand, c->gate_lp, c->sel_word, c->sel_gate,
                                                    out_words, out_gates, rows, t, T);
  VSR_CHECK_CUDA(cudaGetLastError()); c->launches++;
  return VSR_OK;
}

int launch_sample_pick(Ctx* c, int rows, int t, int T, uint64_t seed, int64_t* out_words, int64_t* out_gates,
                       float* lp_words, float* lp_gates, cudaStream_t st) {
  PhaseScope ps(c, PH_BEAM, st);
  k_sample_pick<<<rows, 256, 0, st>>>(c->logits, c->NE, c->V, c->row_max, c->row_lsum, c->gate_lp, (unsigned long long)seed, t, T,
                                      c->sel_word, c->sel_gate, out_words, out_gates, lp_words, lp_gates);
  VSR_CHECK_CUDA(cudaGetLastError()); c->launches++;
  return VSR_OK;
}

int launch_backtrack(Ctx* c, int b, int k, int T, int out_size, int64_t* out_words,
                     int64_t* out_gates, float* lp_words, float* lp_gates, cudaStream_t st) {
  PhaseScope ps(c, PH_FINAL, st);
  VSR_REQUIRE(T <= VSR_MAX_SEQ_LEN, VSR_EINVAL, "seq_len=%d > %d unsupported by the back-track kernel", T, VSR_MAX_SEQ_LEN);
  // final scores come from the history slice of the last step, NOT from the ping-pong buffer c->seq_lp: a replayed
  // CUDA graph bakes in the ping-pong parity of its capture, which need not be the host's current one (odd seq_len)
  k_backtrack<<<b, 64, (size_t)T * k * 20, st>>>(b, k, T, out_size, c->hist_score + (size_t)(T - 1) * b * k, c->hist_parent, c->hist_word,
                                            c->hist_gate, c->hist_lpw, c->hist_lpg, out_words, out_gates,
                                            lp_words, lp_gates);
  VSR_CHECK_CUDA(cudaGetLastError()); c->launches++;
  return VSR_OK;
}

}  // namespace vsr
