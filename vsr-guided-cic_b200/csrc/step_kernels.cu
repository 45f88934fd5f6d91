// Per-step kernels of the role-shift decoder: LSTM-cell / gate pointwise math, the slot
// attention + shift-gate kernel, and the fused log-softmax + verb forcing + per-row top-k.
// Reference math: models/controllable_captioning.py:117-190 (step) and :192-297 (step_v);
// algebra and buffer names follow SURVEY.md Appendix A.
#include <float.h>
#include <math.h>

#include <cuda_fp16.h>

#include "common.cuh"

namespace vsr {

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// store x as fp32 and as the error-compensated fp16 pair the tcgen05 GEMMs consume
struct PairOut { __half* hi; __half* lo; };
__device__ __forceinline__ void store_pair(const PairOut& o, size_t i, float v) {
  if (o.hi == nullptr) return;
  v = fminf(fmaxf(v, -65504.f), 65504.f);   // fp16 range (saturate, never inf)
  const __half h = __float2half_rn(v);
  o.hi[i] = h;
  o.lo[i] = __float2half_rn(v - __half2float(h));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------- LSTM cell 1 + sentinel gate
// (FFMA-twin path; the tensor-core path runs this math in the GEMM-A epilogue.)
// pre1 columns are gate-interleaved: gate g of unit u at cell_col(g, u, 6); gates i,f,g,o,s,gq.
// c1' = sig(f) c1 + sig(i) tanh(g); h1' = sig(o) tanh(c1'); s_t = sig(s) tanh(c1')   (:151-154)
__global__ void k_lstm1(const float* __restrict__ pre1, int ld_pre, const float* __restrict__ c1,
                        float* __restrict__ h1n, float* __restrict__ c1n, float* __restrict__ s_t,
                        float* __restrict__ gq, PairOut h1n_b, PairOut s_t_b, int ld, int H, int rows) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (u >= H || n >= rows) return;
  const float* p = pre1 + (size_t)n * ld_pre;
  const float ig = sigmoidf_(p[cell_col(0, u, 6)]), fg = sigmoidf_(p[cell_col(1, u, 6)]),
              gg = tanhf(p[cell_col(2, u, 6)]), og = sigmoidf_(p[cell_col(3, u, 6)]),
              sg = sigmoidf_(p[cell_col(4, u, 6)]);
  gq[(size_t)n * ld + u] = p[cell_col(5, u, 6)];
  const float c = fg * c1[(size_t)n * ld + u] + ig * gg;
  const float tc = tanhf(c);
  c1n[(size_t)n * ld + u] = c;
  const float hv = og * tc, sv = sg * tc;
  h1n[(size_t)n * ld + u] = hv;
  s_t[(size_t)n * ld + u] = sv;
  store_pair(h1n_b, (size_t)n * ld + u, hv);
  store_pair(s_t_b, (size_t)n * ld + u, sv);
}

// g_t = sig(gq + hg) * tanh(c1')   (:181-182; hg = W1_hg . h1' comes from the h1' GEMM)
__global__ void k_gt(const float* __restrict__ gq_in, const float* __restrict__ hb, int ld_hb,
                     const float* __restrict__ c1n, float* __restrict__ g_t, PairOut g_t_b, int ld, int H,
                     int rows) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (u >= H || n >= rows) return;
  const float gq = gq_in[(size_t)n * ld + u] + hb[(size_t)n * ld_hb + u];
  const float gv = sigmoidf_(gq) * tanhf(c1n[(size_t)n * ld + u]);
  g_t[(size_t)n * ld + u] = gv;
  store_pair(g_t_b, (size_t)n * ld + u, gv);
}

// LSTM cell 2 (FFMA-twin path): pre2 columns gate-interleaved, cell_col(g, u, 4)   (:177 / :258)
__global__ void k_lstm2(const float* __restrict__ pre2, int ld_pre, const float* __restrict__ c2,
                        float* __restrict__ h2n, float* __restrict__ c2n, PairOut h2n_b, int ld, int H,
                        int rows) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (u >= H || n >= rows) return;
  const float* p = pre2 + (size_t)n * ld_pre;
  const float ig = sigmoidf_(p[cell_col(0, u, 4)]), fg = sigmoidf_(p[cell_col(1, u, 4)]),
              gg = tanhf(p[cell_col(2, u, 4)]), og = sigmoidf_(p[cell_col(3, u, 4)]);
  const float c = fg * c2[(size_t)n * ld + u] + ig * gg;
  c2n[(size_t)n * ld + u] = c;
  const float hv = og * tanhf(c);
  h2n[(size_t)n * ld + u] = hv;
  store_pair(h2n_b, (size_t)n * ld + u, hv);
}

// ---------------------------------------------------------------- slot attention + shift gate
// One CTA per row.  Reads only the VALID region rows of the row's current slot tile
// det_seqs[caption, ptr] (128-bit coalesced loads) plus their hoisted att_va projections P.
//   e_r   = att_a . tanh(P_r + ha)            (padding rows: P_r = 0)          :161-162
//   e_s   = att_s . tanh(sa + ha)                                              :163-164
//   alpha = softmax([e_s, e_0..e_{R-1}]) * mask ; alpha /= sum(alpha)          :167-169
//   att   = alpha_s * sentinel + sum_r alpha_r * region_r                      :171
//   gate  = log_softmax([att_g . tanh(ga + ha), sum_{valid r} e_r])            :184-188
constexpr int ATT_THREADS = 256;
constexpr int ATT_MAX_R = 64;

struct AttendArgs {
  const float* det_seqs;   // (b, L, R, F)
  const float* P;          // [b*L*R][ldP]
  const uint8_t* seq_valid;  // [b*L*R]
  const int32_t* ptr;      // [rows]
  const float* sent; int ld_sent; int o_sa;   // sentinel (F) at 0 | sa (A) at o_sa
  const float* hb; int ld_hb; int o_ha;       // hg (H) at 0 | ha (A) at o_ha | ...
  const float* ga; int ld_ga;
  const float *v_a, *v_s, *v_g;
  float* att; int ld_att;
  PairOut att_b;
  float* gate_lp;          // [rows][2]
  int rows, cur_beam, L, R, F, A, H, ldP;
};

__global__ void __launch_bounds__(ATT_THREADS) k_attend(const AttendArgs a) {
  extern __shared__ float sm[];
  float* ha = sm;                       // [A]
  float* e = ha + a.A;                  // [R+1] scores, later alpha; index 0 = sentinel
  __shared__ float red[ATT_THREADS / 32];
  __shared__ float s_stay, s_sent_sum, s_pad;
  __shared__ uint8_t s_valid[ATT_MAX_R];
  __shared__ int s_rows[ATT_MAX_R];     // compacted valid region rows and their weights
  __shared__ float s_w[ATT_MAX_R];
  __shared__ int s_nv;

  const int n = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarp = ATT_THREADS / 32;
  const int cap = n / a.cur_beam;
  const int slot = a.ptr[n];
  const size_t tile_row0 = ((size_t)cap * a.L + slot) * a.R;
  const float* tile = a.det_seqs + tile_row0 * a.F;
  const float* Pt = a.P + tile_row0 * a.ldP;
  const float* sent = a.sent + (size_t)n * a.ld_sent;   // sentinel feature row
  const float* sa = sent + a.o_sa;
  const float* gav = a.ga + (size_t)n * a.ld_ga;

  for (int i = tid; i < a.A; i += ATT_THREADS) ha[i] = a.hb[(size_t)n * a.ld_hb + a.o_ha + i];
  if (tid < a.R) s_valid[tid] = a.seq_valid[tile_row0 + tid];
  // sum of the sentinel row (its mask is computed like any region row, :159)
  float ssum = 0.f;
  for (int f = tid * 4; f < a.F; f += ATT_THREADS * 4) {
    const float4 v = *reinterpret_cast<const float4*>(sent + f);
    ssum += (v.x + v.y) + (v.z + v.w);
  }
  ssum = warp_sum(ssum);
  if (lane == 0) red[warp] = ssum;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < nwarp; ++w) s += red[w];
    s_sent_sum = s;
  }

  // scores: job 0 = sentinel, 1 = stay-gate logit, 2 = padding-row score, 3.. = valid regions
  for (int job = warp; job < a.R + 3; job += nwarp) {
    const float* add = nullptr;
    const float* vec = a.v_a;
    bool run = true;
    if (job == 0) { add = sa; vec = a.v_s; }
    else if (job == 1) { add = gav; vec = a.v_g; }
    else if (job == 2) { add = nullptr; }
    else { const int r = job - 3; run = s_valid[r] != 0; add = Pt + (size_t)r * a.ldP; }
    if (!run) continue;
    float acc = 0.f;
    for (int i = lane; i < a.A; i += 32) acc += vec[i] * tanhf((add != nullptr ? add[i] : 0.f) + ha[i]);
    acc = warp_sum(acc);
    if (lane == 0) {
      if (job == 0) e[0] = acc;
      else if (job == 1) s_stay = acc;
      else if (job == 2) s_pad = acc;       // score shared by all padding rows
      else e[job - 2] = acc;
    }
  }
  __syncthreads();
  if (tid == 0) {
    const float e_pad = s_pad;
    float m = e[0];
    float shift = 0.f;
    for (int r = 0; r < a.R; ++r) {
      if (!s_valid[r]) e[r + 1] = e_pad; else shift += e[r + 1];
      m = fmaxf(m, e[r + 1]);
    }
    float S = 0.f;
    for (int r = 0; r <= a.R; ++r) { e[r] = expf(e[r] - m); S += e[r]; }
    float T = 0.f;
    for (int r = 0; r <= a.R; ++r) {
      const bool valid = (r == 0) ? (s_sent_sum != 0.f) : (s_valid[r - 1] != 0);
      e[r] = valid ? e[r] / S : 0.f;
      T += e[r];
    }
    int nv = 0;
    for (int r = 0; r <= a.R; ++r) {
      e[r] = e[r] / T;
      if (r > 0 && s_valid[r - 1]) { s_rows[nv] = r - 1; s_w[nv] = e[r]; ++nv; }
    }
    s_nv = nv;
    // shift-gate head
    const float stay = s_stay;
    const float gm = fmaxf(stay, shift);
    const float ls = logf(expf(stay - gm) + expf(shift - gm));
    a.gate_lp[(size_t)n * 2 + 0] = (stay - gm) - ls;
    a.gate_lp[(size_t)n * 2 + 1] = (shift - gm) - ls;
  }
  __syncthreads();

  // weighted sum over sentinel + valid regions
  float* out = a.att + (size_t)n * a.ld_att;
  const float a_s = e[0];
  const int nv = s_nv;
  for (int f = tid * 4; f < a.F; f += ATT_THREADS * 4) {
    const float4 sv = *reinterpret_cast<const float4*>(sent + f);
    float4 acc = make_float4(a_s * sv.x, a_s * sv.y, a_s * sv.z, a_s * sv.w);
    int i = 0;
    for (; i + 4 <= nv; i += 4) {   // four independent 128-bit loads in flight per thread
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = __ldg(reinterpret_cast<const float4*>(tile + (size_t)s_rows[i + u] * a.F + f));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float w = s_w[i + u];
        acc.x += w * v[u].x; acc.y += w * v[u].y; acc.z += w * v[u].z; acc.w += w * v[u].w;
      }
    }
    for (; i < nv; ++i) {
      const float w = s_w[i];
      const float4 v = __ldg(reinterpret_cast<const float4*>(tile + (size_t)s_rows[i] * a.F + f));
      acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
    }
    *reinterpret_cast<float4*>(out + f) = acc;
    const size_t o = (size_t)n * a.ld_att + f;
    store_pair(a.att_b, o, acc.x); store_pair(a.att_b, o + 1, acc.y);
    store_pair(a.att_b, o + 2, acc.z); store_pair(a.att_b, o + 3, acc.w);
  }
}

// ---------------------------------------------------------------- log-softmax + verb forcing + top-k
constexpr int SM_THREADS = 256;

__device__ __forceinline__ int64_t load_verb(const void* verbs, int dtype, size_t i) {
  if (dtype == VSR_DT_F64) return (int64_t) reinterpret_cast<const double*>(verbs)[i];
  if (dtype == VSR_DT_F32) return (int64_t) reinterpret_cast<const float*>(verbs)[i];
  return reinterpret_cast<const int64_t*>(verbs)[i];
}

// (value desc, index asc) total order: is (v1,i1) before (v2,i2)?
__device__ __forceinline__ bool before(float v1, int i1, float v2, int i2) {
  return v1 > v2 || (v1 == v2 && i1 < i2);
}

struct SoftmaxArgs {
  const float* logits; int ld;   // [rows][ld]
  int rows, V, cur_beam, L, topk;
  const int32_t* ptr;
  const void* verbs; int verbs_dtype; int use_verbs, gt;
  const int64_t* vt_keys; const int32_t* vt_off; const int32_t* vt_idx; int vt_n;
  float* row_max; float* row_lsum; int32_t* forced; int32_t* cand;  // cand [rows][VSR_MAX_BEAM]
  float* gate_lp;                 // [rows][2], overwritten with [-1e3, 0] on verb rows
  float* out_logp; int64_t out_stride;   // optional full rows
  float* gate_out; int64_t gate_stride;  // optional copy of the post-forcing gate rows
};

__global__ void __launch_bounds__(SM_THREADS) k_softmax_topk(const SoftmaxArgs a) {
  __shared__ float red_v[SM_THREADS / 32];
  __shared__ int red_i[SM_THREADS / 32];
  __shared__ float s_max, s_lsum;
  __shared__ int s_forced, s_pick;
  const int n = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = SM_THREADS / 32;
  const float* x = a.logits + (size_t)n * a.ld;
  const int V = a.V;

  // pass 1: running max and the thread-local top-K list (K = VSR_MAX_BEAM, registers)
  float tv[VSR_MAX_BEAM]; int ti[VSR_MAX_BEAM];
#pragma unroll
  for (int j = 0; j < VSR_MAX_BEAM; ++j) { tv[j] = -INFINITY; ti[j] = 0x7fffffff; }
  float mx = -INFINITY;
  for (int v0 = tid * 4; v0 < V; v0 += SM_THREADS * 4) {
    const float4 q = *reinterpret_cast<const float4*>(x + v0);
    const float qs[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int v = v0 + u;
      if (v < V) {
        const float val = qs[u];
        mx = fmaxf(mx, val);
        if (before(val, v, tv[VSR_MAX_BEAM - 1], ti[VSR_MAX_BEAM - 1])) {
          float cv = val; int ci = v;   // insertion into the sorted list
#pragma unroll
          for (int j = 0; j < VSR_MAX_BEAM; ++j) {
            if (before(cv, ci, tv[j], ti[j])) {
              const float t1 = tv[j]; const int t2 = ti[j];
              tv[j] = cv; ti[j] = ci; cv = t1; ci = t2;
            }
          }
        }
      }
    }
  }
  mx = warp_max(mx);
  if (lane == 0) red_v[warp] = mx;
  __syncthreads();
  if (tid == 0) {
    float m = red_v[0];
    for (int w = 1; w < NW; ++w) m = fmaxf(m, red_v[w]);
    s_max = m;
  }
  __syncthreads();
  mx = s_max;
  // pass 2: sum exp(x - max)
  float se = 0.f;
  for (int v0 = tid * 4; v0 < V; v0 += SM_THREADS * 4) {
    const float4 q = *reinterpret_cast<const float4*>(x + v0);
    se += (v0 + 0 < V ? expf(q.x - mx) : 0.f) + (v0 + 1 < V ? expf(q.y - mx) : 0.f) +
          (v0 + 2 < V ? expf(q.z - mx) : 0.f) + (v0 + 3 < V ? expf(q.w - mx) : 0.f);
  }
  se = warp_sum(se);
  __syncthreads();
  if (lane == 0) red_v[warp] = se;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < NW; ++w) s += red_v[w];
    const float lsum = logf(s);
    s_lsum = lsum;
    // verb forcing (:271-295): which vocabulary index does the current slot force, if any
    int forced = -1;
    if (a.use_verbs && a.verbs != nullptr) {
      const int cap = n / a.cur_beam;
      const int64_t verb = load_verb(a.verbs, a.verbs_dtype, (size_t)cap * a.L + a.ptr[n]);
      if (verb != -1) {
        if (a.gt) {
          forced = (int)verb;
        } else {
          forced = 0;   // key missing or empty list -> vocabulary index 0 (:291-292)
          int lo = 0, hi = a.vt_n - 1, pos = -1;
          while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const int64_t k = a.vt_keys[mid];
            if (k == verb) { pos = mid; break; }
            if (k < verb) lo = mid + 1; else hi = mid - 1;
          }
          if (pos >= 0 && a.vt_off[pos + 1] > a.vt_off[pos]) {
            float best = -1e6f; int best_i = -1;    // strict '>' : first maximum wins (:284-289)
            for (int q = a.vt_off[pos]; q < a.vt_off[pos + 1]; ++q) {
              const int idx = a.vt_idx[q];
              const float lp = (x[idx] - mx) - lsum;
              if (lp > best) { best = lp; best_i = idx; }
            }
            forced = best_i < 0 ? V - 1 : best_i;   // python index -1 == last vocabulary entry
          }
        }
        forced = min(max(forced, 0), V - 1);
      }
    }
    s_forced = forced;
    a.row_max[n] = mx;
    a.row_lsum[n] = lsum;
    a.forced[n] = forced;
    if (forced >= 0) { a.gate_lp[(size_t)n * 2] = -1e3f; a.gate_lp[(size_t)n * 2 + 1] = 0.f; }
    if (a.gate_out != nullptr) {
      a.gate_out[(size_t)n * a.gate_stride + 0] = a.gate_lp[(size_t)n * 2 + 0];
      a.gate_out[(size_t)n * a.gate_stride + 1] = a.gate_lp[(size_t)n * 2 + 1];
    }
  }
  __syncthreads();
  const float lsum = s_lsum;
  const int forced = s_forced;

  if (a.out_logp != nullptr) {
    float* o = a.out_logp + (size_t)n * a.out_stride;
    for (int v = tid; v < V; v += SM_THREADS)
      o[v] = forced >= 0 ? (v == forced ? 0.f : -1e6f) : (x[v] - mx) - lsum;
  }
  if (a.topk <= 0) return;

  int32_t* cd = a.cand + (size_t)n * VSR_MAX_BEAM;
  if (forced >= 0) {
    // forced word first, then the lowest other indices (all tied at -1e6)
    if (tid == 0) {
      cd[0] = forced;
      int v = 0;
      for (int j = 1; j < a.topk; ++j) { if (v == forced) ++v; cd[j] = min(v, V - 1); ++v; }
    }
    return;
  }
  // merge the thread-local lists: topk rounds of block arg-best over list heads
  int head = 0;
  for (int j = 0; j < a.topk; ++j) {
    float hv = -INFINITY; int hi = 0x7fffffff;
#pragma unroll
    for (int q = 0; q < VSR_MAX_BEAM; ++q) if (q == head) { hv = tv[q]; hi = ti[q]; }
    float bv = hv; int bi = hi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (before(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      float fv = red_v[0]; int fi = red_i[0];
      for (int w = 1; w < NW; ++w) if (before(red_v[w], red_i[w], fv, fi)) { fv = red_v[w]; fi = red_i[w]; }
      s_pick = fi;
      cd[j] = fi;
    }
    __syncthreads();
    if (hi == s_pick && head < VSR_MAX_BEAM) ++head;   // the owner pops its head
    __syncthreads();
  }
}

}  // namespace

// ---------------------------------------------------------------- one decoder step (host side)
static PairOut pair_out(const Ctx* c, const F16Pair& b) {
  PairOut o;
  o.hi = c->use_tc ? (__half*)b.hi : nullptr;
  o.lo = c->use_tc ? (__half*)b.lo : nullptr;
  return o;
}

int run_step(Ctx* c, const StepIO& io, cudaStream_t st) {
  const int rows = io.rows, H = c->H;
  VSR_REQUIRE(rows > 0 && rows <= c->cap_rows, VSR_ESTATE, "run_step: rows=%d cap=%d", rows, c->cap_rows);
  VSR_REQUIRE(c->R <= ATT_MAX_R, VSR_EINVAL, "max_detections per slot R=%d > %d", c->R, ATT_MAX_R);
  const bool h2f = c->d.h2_first_lstm != 0;
  const dim3 pw_grid((H + 127) / 128, rows);
  bool fused = false;   // tensor-core path: cell / gate pointwise math runs inside the GEMM epilogues

  {  // A: pre1 = [h2 | xt | h1_old] . WA^T + U[img]
    PhaseScope ps(c, PH_GEMM_A, st);
    GemmArgs g{};
    int s = 0;
    if (h2f) g.seg[s++] = {c->h2, c->Hp, c->Hp, c->Hp, &c->h2_b};
    g.seg[s++] = {c->xt, c->Ep, c->Ep, c->Ep, &c->xt_b};
    g.seg[s++] = {c->h1, c->Hp, c->Hp, c->Hp, &c->h1_b};
    g.nseg = s;
    g.w = c->WA; g.ldw = c->KA; g.wb = &c->WA_b;
    g.rowadd = c->U; g.ld_rowadd = c->NA; g.row_div = io.cur_beam; g.rowadd_mul = (c->n_img == 1 ? 0 : 1);
    g.c = c->pre1; g.ldc = c->NA; g.M = rows; g.N = c->NA;
    fused = gemm_uses_tc(c, g);
    if (fused) {   // LSTM cell 1 + sentinel gate in the epilogue: pre1 is never written
      g.cell.mode = 1; g.cell.c_old = c->c1; g.cell.c_new = c->c1n; g.cell.h_new = c->h1n;
      g.cell.h_hi = c->h1n_b.hi; g.cell.h_lo = c->h1n_b.lo;
      g.cell.s_new = c->s_t; g.cell.s_hi = c->s_t_b.hi; g.cell.s_lo = c->s_t_b.lo;
      g.cell.gq = c->gq; g.cell.ld_state = c->Hp;
    }
    VSR_TRY(launch_gemm(c, g, st)); c->launches++;
  }
  if (!fused) {
    PhaseScope ps(c, PH_LSTM1, st);
    k_lstm1<<<pw_grid, 128, 0, st>>>(c->pre1, c->NA, c->c1, c->h1n, c->c1n, c->s_t, c->gq, pair_out(c, c->h1n_b),
                                     pair_out(c, c->s_t_b), c->Hp, H, rows);
    VSR_CHECK_CUDA(cudaGetLastError()); c->launches++;
  }
  {  // B: [sentinel | sa] = WB1 . s_t + b ;  [hg | ha | pre2_h1] = WB2 . h1'
    PhaseScope ps(c, PH_GEMM_B, st);
    GemmArgs g{};
    g.nseg = 1; g.seg[0] = {c->s_t, c->Hp, c->Hp, c->Hp, &c->s_t_b};
    g.w = c->WB1; g.ldw = c->Hp; g.bias = c->bB1; g.wb = &c->WB1_b;
    g.c = c->sent; g.ldc = c->NB1; g.M = rows; g.N = c->NB1;
    GemmArgs g2{};
    g2.nseg = 1; g2.seg[0] = {c->h1n, c->Hp, c->Hp, c->Hp, &c->h1n_b};
    g2.w = c->WB2; g2.ldw = c->Hp; g2.wb = &c->WB2_b;
    g2.c = c->hb; g2.ldc = c->NB2; g2.M = rows; g2.N = c->NB2;
    if (fused) {   // g_t = sig(gq + W1_hg.h1') * tanh(c1') on the hg column block of the h1' projection
      g2.cell.gt_cols = c->oB2_ha; g2.cell.gt_gq = c->gq; g2.cell.gt_c1n = c->c1n; g2.cell.g_t = c->g_t;
      g2.cell.g_hi = c->g_t_b.hi; g2.cell.g_lo = c->g_t_b.lo; g2.cell.ld_state = c->Hp;
    }
    VSR_TRY(launch_gemm(c, g, st, &g2)); c->launches += fused ? 1 : 2;   // one grouped launch on tensor cores
  }
  if (!fused) {
    PhaseScope ps(c, PH_GT, st);
    k_gt<<<pw_grid, 128, 0, st>>>(c->gq, c->hb, c->NB2, c->c1n, c->g_t, pair_out(c, c->g_t_b), c->Hp, H, rows);
    VSR_CHECK_CUDA(cudaGetLastError()); c->launches++;
  }
  {  // C: ga = att_ga . g_t
    PhaseScope ps(c, PH_GEMM_C, st);
    GemmArgs g{};
    g.nseg = 1; g.seg[0] = {c->g_t, c->Hp, c->Hp, c->Hp, &c->g_t_b};
    g.w = c->WC; g.ldw = c->Hp; g.wb = &c->WC_b;
    g.c = c->ga; g.ldc = c->NC; g.M = rows; g.N = c->NC;
    VSR_TRY(launch_gemm(c, g, st)); c->launches++;
  }
  {
    PhaseScope ps(c, PH_ATTEND, st);
    AttendArgs a{};
    a.det_seqs = c->det_seqs; a.P = c->P; a.seq_valid = c->seq_valid; a.ptr = c->ptr;
    a.sent = c->sent; a.ld_sent = c->NB1; a.o_sa = c->oB1_sa;
    a.hb = c->hb; a.ld_hb = c->NB2; a.o_ha = c->oB2_ha; a.ga = c->ga; a.ld_ga = c->NC;
    a.v_a = c->v_a; a.v_s = c->v_s; a.v_g = c->v_g;
    a.att = c->att; a.ld_att = c->Fp; a.att_b = pair_out(c, c->att_b); a.gate_lp = c->gate_lp;
    a.rows = rows; a.cur_beam = io.cur_beam; a.L = c->L; a.R = c->R; a.F = c->F; a.A = c->A; a.H = H;
    a.ldP = c->NVA;
    const size_t smem = sizeof(float) * (size_t)(c->A + c->R + 1);
    k_attend<<<rows, ATT_THREADS, smem, st>>>(a);
    VSR_CHECK_CUDA(cudaGetLastError()); c->launches++;
  }
  {  // D: pre2 = pre2_h1 + WD . [att | h2_old] + b (+ U2[img])
    PhaseScope ps(c, PH_GEMM_D, st);
    GemmArgs g{};
    g.nseg = 2;
    g.seg[0] = {c->att, c->Fp, c->Fp, c->Fp, &c->att_b};
    g.seg[1] = {c->h2, c->Hp, c->Hp, c->Hp, &c->h2_b};
    g.w = c->WD; g.ldw = c->KD; g.bias = c->bD; g.wb = &c->WD_b;
    g.cadd = c->hb + c->oB2_p2; g.ld_cadd = c->NB2;
    if (c->d.img_second_lstm) {
      g.rowadd = c->U2; g.ld_rowadd = c->ND; g.row_div = io.cur_beam; g.rowadd_mul = (c->n_img == 1 ? 0 : 1);
    }
    g.c = c->pre2; g.ldc = c->ND; g.M = rows; g.N = c->ND;
    if (fused) {   // LSTM cell 2 in the epilogue
      g.cell.mode = 2; g.cell.c_old = c->c2; g.cell.c_new = c->c2n; g.cell.h_new = c->h2n;
      g.cell.h_hi = c->h2n_b.hi; g.cell.h_lo = c->h2n_b.lo; g.cell.ld_state = c->Hp;
    }
    VSR_TRY(launch_gemm(c, g, st)); c->launches++;
  }
  if (!fused) {
    PhaseScope ps(c, PH_LSTM2, st);
    k_lstm2<<<pw_grid, 128, 0, st>>>(c->pre2, c->ND, c->c2, c->h2n, c->c2n, pair_out(c, c->h2n_b), c->Hp, H, rows);
    VSR_CHECK_CUDA(cudaGetLastError()); c->launches++;
  }
  {  // E: logits = out_fc . h2' + b
    PhaseScope ps(c, PH_GEMM_E, st);
    GemmArgs g{};
    g.nseg = 1; g.seg[0] = {c->h2n, c->Hp, c->Hp, c->Hp, &c->h2n_b};
    g.w = c->WE; g.ldw = c->Hp; g.bias = c->bE; g.wb = &c->WE_b;
    g.c = c->logits; g.ldc = c->NE; g.M = rows; g.N = c->NE;
    VSR_TRY(launch_gemm(c, g, st)); c->launches++;
  }
  {
    PhaseScope ps(c, PH_SOFTMAX_TOPK, st);
    SoftmaxArgs a{};
    a.logits = c->logits; a.ld = c->NE; a.rows = rows; a.V = c->V; a.cur_beam = io.cur_beam; a.L = c->L;
    a.topk = io.topk; a.ptr = c->ptr;
    a.verbs = c->verbs; a.verbs_dtype = c->verbs_dtype; a.use_verbs = io.use_verbs; a.gt = io.gt;
    a.vt_keys = c->vt_keys; a.vt_off = c->vt_off; a.vt_idx = c->vt_idx; a.vt_n = c->vt_n;
    a.row_max = c->row_max; a.row_lsum = c->row_lsum; a.forced = c->forced; a.cand = c->cand;
    a.gate_lp = c->gate_lp; a.out_logp = io.out_logp; a.out_stride = io.out_stride;
    a.gate_out = io.gate_out; a.gate_stride = io.gate_stride;
    k_softmax_topk<<<rows, SM_THREADS, 0, st>>>(a);
    VSR_CHECK_CUDA(cudaGetLastError()); c->launches++;
  }
  return VSR_OK;
}

}  // namespace vsr
