// Host bookkeeping of the eval pre-step: the integer logic of coco_scripts/eval_coco.py:148-237 around the two
// device networks, for a whole batch of captions per call.  Pure host code (no kernel here): at a thousand captions per call
// the Python form of this bookkeeping (vsrdec/preorder.py, kept as the readable statement and the test twin) costs more than
// the device work it feeds.
//
//   vsr_preorder_begin   eval_coco.py:148-171  per (caption, verb): the distinct roles in first-seen order (at most `limit`), the
//                                              slots holding each role, the roles held by several slots
//   vsr_preorder_fill    :170-183              inputs of the two batched device calls: verbs / roles / role counts of the S-level
//                                              problems; for every repeated role the rows of its slots (gather indices, -1 = zero row)
//   vsr_preorder_end     :190-237              region order from the assignment, rank assembly per verb, verb_rank_merge
//                                              (utils/tools.py:35-71) over a caption's verbs, permutation of the slot list
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace vsr {
namespace {

struct PreProblem {
  int caption;
  int64_t verb;
  std::vector<int64_t> roles;               // distinct, first-seen order
  std::vector<std::vector<int>> slots;      // per role: the slots holding it, in the order found
  std::vector<int> repeated;                // indices into roles, first-repeat order
};

struct PreState {
  int C, n_verb, L, limit, N, fixed_len;
  std::vector<PreProblem> problems;
  std::vector<std::pair<int, int>> rep;     // (problem, role index) of every R-level problem
};

// utils/tools.py:35-71
std::vector<int> merge_ranks(const std::vector<int>& first, std::vector<int> second) {
  auto in = [](const std::vector<int>& v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); };
  std::vector<int> shared, where;
  for (int s : first) {
    auto it = std::find(second.begin(), second.end(), s);
    if (it != second.end()) { shared.push_back(s); where.push_back((int)(it - second.begin())); }
  }
  std::vector<int> sorted_where = where;
  std::sort(sorted_where.begin(), sorted_where.end());
  if (where != sorted_where)
    for (size_t j = 0; j < shared.size(); ++j) second[sorted_where[j]] = shared[j];
  // nearest shared slot to the right of every slot only the second list has (a later duplicate overwrites, as the dict does)
  std::vector<std::pair<int, int>> anchor;     // (slot, anchor or -1)
  int nearest = -1;
  for (auto it = second.rbegin(); it != second.rend(); ++it) {
    if (in(shared, *it)) { nearest = *it; continue; }
    bool found = false;
    for (auto& a : anchor) if (a.first == *it) { a.second = nearest; found = true; }
    if (!found) anchor.push_back({*it, nearest});
  }
  std::vector<int> merged = first;
  for (int s : second) {
    if (in(shared, s)) continue;
    int a = -1;
    for (auto& p : anchor) if (p.first == s) a = p.second;
    if (a < 0) { merged.push_back(s); continue; }
    auto it = std::find(merged.begin(), merged.end(), a);
    if (it != merged.end()) merged.insert(it, s);
  }
  return merged;
}

}  // namespace
}  // namespace vsr

using vsr::PreProblem;
using vsr::PreState;

extern "C" {

int vsr_preorder_begin(const int64_t* control_verb, const int64_t* det_seqs_v, const int64_t* det_seqs_sr, int32_t C, int32_t n_verb,
                       int32_t L, int32_t limit, int32_t N, int32_t fixed_len, vsr_preorder_handle* out, int32_t* n_problems,
                       int32_t* n_repeated, int32_t* max_roles) {
  if (!control_verb || !det_seqs_v || !det_seqs_sr || !out || !n_problems || !n_repeated || !max_roles) {
    vsr::set_error("vsr_preorder_begin: null argument");
    return VSR_EINVAL;
  }
  VSR_REQUIRE(C >= 0 && n_verb >= 1 && L >= 1 && limit >= 1 && N >= 1 && fixed_len >= 1, VSR_EINVAL, "vsr_preorder_begin: bad sizes");
  PreState* st = new PreState{C, n_verb, L, limit, N, fixed_len, {}, {}};
  int mr = 0;
  for (int c = 0; c < C; ++c) {
    const int64_t* v = det_seqs_v + (size_t)c * L * n_verb;
    const int64_t* sr = det_seqs_sr + (size_t)c * L * n_verb;
    for (int vi = 0; vi < n_verb; ++vi) {
      const int64_t verb = control_verb[(size_t)c * n_verb + vi];
      if (verb == 0) break;                                   // eval_coco.py:150-151
      PreProblem p{c, verb, {}, {}, {}};
      for (int j = 0; j < L; ++j)
        for (int k = 0; k < n_verb; ++k) {
          if (v[j * n_verb + k] != verb || (int)p.roles.size() >= limit) continue;     // :156 (`find_sr < 10` guards both branches)
          const int64_t role = sr[j * n_verb + k];
          const auto it = std::find(p.roles.begin(), p.roles.end(), role);
          if (it == p.roles.end()) { p.roles.push_back(role); p.slots.push_back({j}); }
          else {
            const int ri = (int)(it - p.roles.begin());
            p.slots[ri].push_back(j);
            if (std::find(p.repeated.begin(), p.repeated.end(), ri) == p.repeated.end()) p.repeated.push_back(ri);
          }
        }
      if (p.roles.empty()) continue;                          // :170-171
      mr = std::max(mr, (int)p.roles.size());
      const int pi = (int)st->problems.size();
      for (int ri : p.repeated) st->rep.push_back({pi, ri});
      st->problems.push_back(std::move(p));
    }
  }
  *out = (vsr_preorder_handle)st;
  *n_problems = (int32_t)st->problems.size();
  *n_repeated = (int32_t)st->rep.size();
  *max_roles = mr;
  return VSR_OK;
}

int vsr_preorder_fill(vsr_preorder_handle h, int64_t* verbs, int64_t* roles, int32_t roles_ld, int32_t* counts, int64_t* gather) {
  if (!h) { vsr::set_error("vsr_preorder_fill: null handle"); return VSR_EINVAL; }
  PreState* st = (PreState*)h;
  VSR_REQUIRE(roles_ld >= st->limit, VSR_EINVAL, "vsr_preorder_fill: roles_ld=%d < limit=%d", roles_ld, st->limit);
  for (size_t i = 0; i < st->problems.size(); ++i) {
    const PreProblem& p = st->problems[i];
    if (verbs) verbs[i] = p.verb;
    if (counts) counts[i] = (int32_t)p.roles.size();
    if (roles) {
      for (int k = 0; k < roles_ld; ++k) roles[i * roles_ld + k] = k < (int)p.roles.size() ? p.roles[k] : 0;
    }
  }
  if (gather)
    for (size_t n = 0; n < st->rep.size(); ++n) {
      const PreProblem& p = st->problems[st->rep[n].first];
      const std::vector<int>& locs = p.slots[st->rep[n].second];
      for (int j = 0; j < st->N; ++j)
        gather[n * st->N + j] = j < (int)locs.size() ? (int64_t)p.caption * st->fixed_len + locs[j] : -1;     // :178-182
    }
  return VSR_OK;
}

void vsr_preorder_free(vsr_preorder_handle h) { delete (PreState*)h; }

int vsr_preorder_end(vsr_preorder_handle h, const int64_t* pred, int32_t pred_ld, const int32_t* assign, const uint8_t* slot_valid,
                     const double* verb_list, int64_t* src_slot, float* verbs_out) {
  if (!h) { vsr::set_error("vsr_preorder_end: null handle"); return VSR_EINVAL; }
  PreState* st = (PreState*)h;
  const int C = st->C, FL = st->fixed_len, N = st->N;
  if ((!st->problems.empty() && !pred) || (!st->rep.empty() && !assign) || !slot_valid || !verb_list || !src_slot || !verbs_out) {
    vsr::set_error("vsr_preorder_end: null argument");
    return VSR_EINVAL;
  }
  // R level: slots of a repeated role in the order of their assigned columns (eval_coco.py:190-200)
  std::vector<std::vector<int>> region(st->rep.size());
  for (size_t n = 0; n < st->rep.size(); ++n) {
    const std::vector<int>& locs = st->problems[st->rep[n].first].slots[st->rep[n].second];
    const int m = std::min((int)locs.size(), N);
    std::vector<int> order(m);
    for (int a = 0; a < m; ++a) order[a] = a;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return assign[n * N + a] < assign[n * N + b]; });
    for (int a : order) region[n].push_back(locs[a]);
  }
  // rank per problem, merged over a caption's verbs (:202-215)
  std::vector<std::vector<int>> final_rank(C);
  std::vector<char> started(C, 0);
  size_t rep_at = 0;
  for (size_t i = 0; i < st->problems.size(); ++i) {
    const PreProblem& p = st->problems[i];
    const size_t rep0 = rep_at;                         // this problem's R-level problems are rep[rep0, rep0 + repeated.size())
    rep_at += p.repeated.size();
    std::vector<int> rank;
    for (int t = 0; t < pred_ld; ++t) {
      const int64_t role = pred[i * pred_ld + t];
      if (role == 0) break;
      const auto it = std::find(p.roles.begin(), p.roles.end(), role);
      if (it == p.roles.end()) continue;
      const int ri = (int)(it - p.roles.begin());
      if (p.slots[ri].size() != 1) {
        const auto r = std::find(p.repeated.begin(), p.repeated.end(), ri);
        const std::vector<int>& reg = region[rep0 + (size_t)(r - p.repeated.begin())];
        rank.insert(rank.end(), reg.begin(), reg.end());
      } else {
        rank.push_back(p.slots[ri][0]);
      }
    }
    if (!started[p.caption]) { final_rank[p.caption] = rank; started[p.caption] = 1; }
    else final_rank[p.caption] = vsr::merge_ranks(final_rank[p.caption], rank);
  }
  // permutation of the slot list (:217-237)
  for (int c = 0; c < C; ++c) {
    const std::vector<int>& fr = final_rank[c];
    const int placed = std::min((int)fr.size(), FL);
    int n_kept = 0, last = -1;
    for (int j = 0; j < placed; ++j) {
      const int r = fr[j];
      if (r >= 0 && r < FL && slot_valid[(size_t)c * FL + r]) { src_slot[(size_t)c * FL + n_kept++] = r; last = r; }
    }
    for (int j = n_kept; j < FL; ++j) src_slot[(size_t)c * FL + j] = last;
    for (int j = 0; j < FL; ++j)
      verbs_out[(size_t)c * FL + j] = (j < placed && fr[j] >= 0 && fr[j] < FL) ? (float)verb_list[(size_t)c * FL + fr[j]] : -1.f;
  }
  delete st;
  return VSR_OK;
}

}  // extern "C"
