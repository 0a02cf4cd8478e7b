// FeatureBank.update for all objects in one host call (FeatureBank.py:53-115): the sequence
//   prepare candidates -> match -> plan -> merge -> [sync: |A|] -> (evict plan -> [sync] -> compaction) -> append -> clamp
// issued from C++.  The arithmetic lives in the kernels of vfn_bank.cu / vfn_tc.cu / vfn_simt.cu; this file only
// orders the launches, so a frame costs one ctypes call instead of ~25 (profiles/r1a_summary.md: the Python launch
// path was 38 % of the frame).
#include "vfn_common.cuh"
#include "vfn_tc.cuh"

namespace vfn {

struct UpdLayout {
  size_t ck, cv, nck, ncv;     // per-object strides apply: base + obj * stride
  size_t s_ck, s_cv;
  size_t counts, plan, lfu, s_lfu, pws, s_pws, cws, s_cws, match, total;
};

static UpdLayout upd_layout(int obj_n, int64_t n_max, int64_t hw, int d_key, int d_val) {
  UpdLayout L{};
  size_t o = 0;
  L.s_ck = align_up((size_t)hw * d_key * sizeof(float), 256);
  L.s_cv = align_up((size_t)hw * d_val * sizeof(float), 256);
  L.ck = o;  o += obj_n * L.s_ck;
  L.nck = o; o += obj_n * L.s_ck;
  L.cv = o;  o += obj_n * L.s_cv;
  L.ncv = o; o += obj_n * L.s_cv;
  L.counts = o; o += align_up((size_t)obj_n * 4 * sizeof(int32_t), 256);
  L.plan = o;   o += align_up((size_t)obj_n * 72 * sizeof(int32_t), 256);
  L.s_lfu = align_up((size_t)(n_max > 0 ? n_max : 1) * sizeof(float), 256);
  L.lfu = o;    o += obj_n * L.s_lfu;
  L.s_pws = align_up(vfn_bank_plan_workspace_bytes(hw), 256);
  L.pws = o;    o += obj_n * L.s_pws;
  L.s_cws = align_up(vfn_bank_compact_workspace_bytes(n_max > 0 ? n_max : 1), 256);
  L.cws = o;    o += obj_n * L.s_cws;
  L.match = o;
  size_t m_tc = tc_match_workspace_bytes(obj_n, hw);
  size_t m_simt = vfn_bank_match_workspace_bytes(n_max, hw);
  o += align_up(m_tc > m_simt ? m_tc : m_simt, 256);
  L.total = o;
  return L;
}

}  // namespace vfn

using namespace vfn;

extern "C" {

size_t vfn_bank_update_workspace_bytes(int32_t obj_n, int64_t n_max, int64_t hw, int32_t d_key, int32_t d_val) {
  if (obj_n < 1 || hw < 1 || d_key < 1 || d_val < 1) return 0;
  return upd_layout(obj_n, n_max, hw, d_key, d_val).total;
}

int vfn_bank_update(vfn_bank* banks, vfn_bank* alts, int32_t obj_n, vfn_update_io* io, int64_t hw, float frame_idx,
                    float update_rate, float thres_close, double class_budget, void* d_ws, size_t ws_bytes,
                    int32_t* h_pinned, int32_t impl, void* defer_event, void* stream) {
  VFN_CHECK_ARG(banks && alts && io && d_ws && h_pinned && obj_n >= 1 && obj_n <= 4 && hw >= 1, "bank_update: bad args");
  const int d_key = banks[0].d_key, d_val = banks[0].d_val;
  int64_t n_max = 0;
  for (int c = 0; c < obj_n; ++c) {
    VFN_CHECK_ARG(banks[c].d_key == d_key && banks[c].d_val == d_val, "bank_update: banks differ in dims");
    VFN_CHECK_ARG(banks[c].n >= 1 && banks[c].n + hw <= banks[c].cap, "bank_update: bank %d needs cap >= n + hw", c);
    VFN_CHECK_ARG(io[c].d_prev_key_dm && io[c].d_prev_value_dm && io[c].d_match_idx && io[c].d_match_corr &&
                      io[c].d_merge_q && io[c].d_merge_slot && io[c].d_run_off && io[c].d_append_q,
                  "bank_update: io[%d] has NULL buffers", c);
    VFN_CHECK_ARG(!banks[c].n_live || (banks[c].n_min >= 1 && banks[c].n_min <= banks[c].n),
                  "bank_update: bank %d has n_live but n_min=%lld is not a lower bound of n=%lld", c,
                  (long long)banks[c].n_min, (long long)banks[c].n);
    if (banks[c].n > n_max) n_max = banks[c].n;
    io[c].n_before = banks[c].n;
    io[c].evicted = io[c].swapped = io[c].evict_status = io[c].kept = io[c].n_iter = io[c].deferred = 0;
  }
  const UpdLayout L = upd_layout(obj_n, n_max, hw, d_key, d_val);
  if (ws_bytes < L.total) { set_error("bank_update: workspace %zu < %zu", ws_bytes, L.total); return VFN_E_CAPACITY; }
  cudaStream_t st = as_stream(stream);
  char* ws = reinterpret_cast<char*>(d_ws);
  // entry-major candidates (prev_layout == 1) ARE the raw entry-major rows the merge / append kernels read: no copy
  auto CK = [&](int c) {
    return io[c].prev_layout == 1 ? const_cast<float*>(io[c].d_prev_key_dm) : reinterpret_cast<float*>(ws + L.ck + c * L.s_ck);
  };
  auto NCK = [&](int c) { return reinterpret_cast<float*>(ws + L.nck + c * L.s_ck); };
  auto CV = [&](int c) {
    return io[c].prev_layout == 1 ? const_cast<float*>(io[c].d_prev_value_dm) : reinterpret_cast<float*>(ws + L.cv + c * L.s_cv);
  };
  auto NCV = [&](int c) { return reinterpret_cast<float*>(ws + L.ncv + c * L.s_cv); };
  int32_t* counts = reinterpret_cast<int32_t*>(ws + L.counts);
  int32_t* plan = reinterpret_cast<int32_t*>(ws + L.plan);
  char* mws = ws + L.match;
  int32_t* h_counts = h_pinned;                 // obj_n * 4
  int32_t* h_plan = h_pinned + 4 * obj_n;       // obj_n * 72   (caller provides obj_n * 80 ints)

  const bool tc = (impl != 1) && d_key == 128 && banks[0].nkh != nullptr && vfn_device_is_sm100();
  if (impl == 2 && !tc) { set_error("tcgen05 match needs d_key = 128 operands on an sm_100 device"); return VFN_E_UNSUPPORTED; }
  if (!tc)
    for (int c = 0; c < obj_n; ++c)
      VFN_CHECK_ARG(!banks[c].n_live || banks[c].n_min == banks[c].n,
                    "bank_update: the fp32 SIMT kernels need exact bank sizes (bank %d was passed with bounds)", c);

  // (1) candidates: (d, hw) -> entry-major raw + normalised (FeatureBank.py:64,88), + fp16 split of 16 * normalised keys
  PrepJob jobs[8];
  for (int c = 0; c < obj_n; ++c) {
    const int em = io[c].prev_layout == 1;      // (hw, d) candidates as vfn_keyvalue writes them: no transpose
    jobs[2 * c] = PrepJob{io[c].d_prev_key_dm, d_key, hw, em ? nullptr : CK(c), NCK(c),
                          tc ? tc_match_cand_hi(mws, obj_n, hw, c) : nullptr,
                          tc ? tc_match_cand_lo(mws, obj_n, hw, c) : nullptr, NK_SCALE, 1, em,
                          tc ? tc_operand_rows(hw) : 0};       // pad rows of the match operand zeroed by the same launch
    jobs[2 * c + 1] = PrepJob{io[c].d_prev_value_dm, d_val, hw, em ? nullptr : CV(c), NCV(c), nullptr, nullptr, 1.f, 0, em};
  }
  if (int rc = launch_prep(jobs, 2 * obj_n, st)) return rc;

  // (2) match (FeatureBank.py:63-68): all objects in one tensor-core launch, or the fp32 SIMT kernels per object
  if (tc) {
    const float* nck[4]; int32_t* idx[4]; float* corr[4];
    for (int c = 0; c < obj_n; ++c) { nck[c] = NCK(c); idx[c] = io[c].d_match_idx; corr[c] = io[c].d_match_corr; }
    if (int rc = tc_match(banks, obj_n, nck, hw, mws, 1, idx, corr, st)) return rc;
  } else {
    for (int c = 0; c < obj_n; ++c)
      if (int rc = vfn_bank_match(&banks[c], NCK(c), hw, io[c].d_match_idx, io[c].d_match_corr, mws,
                                  ws_bytes - L.match, 1, stream))
        return rc;
  }
  // (3) plan + merge, one launch each for all objects (FeatureBank.py:71-97)
  UpdObj uo[4];
  for (int c = 0; c < obj_n; ++c) {
    uo[c] = UpdObj{};
    uo[c].bank = banks[c];
    uo[c].ck = CK(c); uo[c].cv = CV(c); uo[c].nck = NCK(c); uo[c].ncv = NCV(c);
    uo[c].match_idx = io[c].d_match_idx; uo[c].match_corr = io[c].d_match_corr;
    uo[c].merge_q = io[c].d_merge_q; uo[c].merge_slot = io[c].d_merge_slot; uo[c].run_off = io[c].d_run_off;
    uo[c].append_q = io[c].d_append_q; uo[c].counts = counts + 4 * c; uo[c].h_counts = h_counts + 4 * c;
    uo[c].plan_ws = ws + L.pws + c * L.s_pws;
    uo[c].n_live = banks[c].n_live;
  }
  if (int rc = launch_plan(uo, obj_n, hw, thres_close, st)) return rc;
  // Deferred completion: when no object can reach its budget whatever |A| turns out to be (n + hw <= class_budget,
  // FeatureBank.py:102 cannot fire), nothing on the host depends on the counts: append and clamp take |A| from device
  // memory, the counts land in h_pinned when the plan kernel retires, and the caller learns the new bank sizes from
  // vfn_bank_update_finish() after waiting on `defer_event` - the stream is never drained.
  bool can_defer = defer_event != nullptr;
  for (int c = 0; c < obj_n; ++c) can_defer = can_defer && !(class_budget < (double)(banks[c].n + hw));
  if (can_defer) {
    VFN_CUDA_OK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(defer_event), st));
    if (int rc = launch_merge(uo, obj_n, hw, update_rate, st)) return rc;
    UpdObj ao[4];
    vfn_bank cb[4];
    for (int c = 0; c < obj_n; ++c) {
      ao[c] = uo[c];
      ao[c].sel = io[c].d_append_q; ao[c].n_sel_dev = counts + 4 * c + 2; ao[c].n_sel = hw;   // grid bound; rows = *n_sel_dev
      cb[c] = banks[c];
      cb[c].n = banks[c].n + hw;            // clamp over an upper bound: rows beyond the live count are dead storage
      io[c].deferred = 1;
      io[c].n_merge = io[c].n_runs = io[c].n_append = -1;
    }
    if (int rc = launch_append(ao, obj_n, frame_idx, 0.f, st)) return rc;
    if (int rc = launch_clamp(cb, obj_n, st)) return rc;
    return VFN_OK;
  }
  // synchronous path: the host needs exact bank sizes from here on
  for (int c = 0; c < obj_n; ++c)
    VFN_CHECK_ARG(!banks[c].n_live || banks[c].n_min == banks[c].n,
                  "bank_update: bank %d was passed with bounds (n_min < n) but this update may evict: finish the "
                  "deferred updates first", c);
  if (int rc = launch_merge(uo, obj_n, hw, update_rate, st)) return rc;
  VFN_CUDA_OK(cudaStreamSynchronize(st));                 // |merge|, |runs|, |append| per object (nonzero/unique syncs)
  bool any_evict = false;
  for (int c = 0; c < obj_n; ++c) {
    io[c].n_merge = h_counts[4 * c + 0];
    io[c].n_runs = h_counts[4 * c + 1];
    io[c].n_append = h_counts[4 * c + 2];
    if (class_budget < (double)(banks[c].n + io[c].n_append)) {            // FeatureBank.py:102
      io[c].evicted = 1;
      any_evict = true;
      if (int rc = vfn_bank_evict_plan(&banks[c], frame_idx, class_budget, io[c].n_append, plan + 72 * c, h_plan + 72 * c,
                                       reinterpret_cast<float*>(ws + L.lfu + c * L.s_lfu), stream))
        return rc;
    }
  }
  if (any_evict) {
    VFN_CUDA_OK(cudaStreamSynchronize(st));               // int(LFU.min()) syncs of remove()
    for (int c = 0; c < obj_n; ++c) {
      if (!io[c].evicted) continue;
      const int32_t* hp = h_plan + 72 * c;
      io[c].evict_status = hp[0]; io[c].kept = hp[1]; io[c].n_iter = hp[2];
      for (int k = 0; k < 64; ++k) io[c].thresholds[k] = (k < hp[2]) ? hp[4 + k] : 0;   // first 64 of n_iter
      if (hp[0] != 0) continue;                             // the caller raises like the reference
      VFN_CHECK_ARG(alts[c].keys != nullptr && alts[c].cap >= hp[1] + hw, "bank_update: eviction needs alts[%d] with cap >= kept + hw", c);
      alts[c].n = 0;
      if (int rc = vfn_bank_compact(&banks[c], &alts[c], reinterpret_cast<float*>(ws + L.lfu + c * L.s_lfu), plan + 72 * c,
                                    ws + L.cws + c * L.s_cws, L.s_cws, stream))
        return rc;
      // algorithmic bytes of the compaction: evicted rows read once, kept rows read + written
      vfn_profile_add_work(PROF_COMPACT, 4.0 * (d_key + d_val + 2) * ((double)(banks[c].n - hp[1]) + 2.0 * hp[1]));
      int32_t* nl = banks[c].n_live;                        // the live count belongs to the bank, not to a slab
      vfn_bank tmp = banks[c]; banks[c] = alts[c]; alts[c] = tmp;
      banks[c].n_live = nl; alts[c].n_live = nullptr;
      banks[c].n = hp[1];
      io[c].swapped = 1;
    }
  }
  // (4) append + clamp (FeatureBank.py:105-115), one launch each
  UpdObj ao[4];
  vfn_bank cb[4];
  int64_t commit[4];
  int n_a = 0, n_c = 0;
  for (int c = 0; c < obj_n; ++c) {
    if (io[c].evict_status != 0) {            // the caller raises; keep the device count equal to the host's
      if (int rc = vfn_bank_set_live(&banks[c], banks[c].n, stream)) return rc;
      continue;
    }
    const int64_t n_app = io[c].n_append;
    if (n_app > 0) {
      ao[n_a] = uo[c];
      ao[n_a].bank = banks[c];                // after a possible ping-pong swap
      ao[n_a].bank.n_live = nullptr;          // the host count is exact here (and an eviction changed it)
      ao[n_a].sel = io[c].d_append_q; ao[n_a].n_sel_dev = nullptr; ao[n_a].n_sel = n_app;
      ++n_a;
    }
    banks[c].n += n_app;
    banks[c].n_min = banks[c].n;
    commit[n_c] = banks[c].n;                 // device-resident count := exact host count
    cb[n_c++] = banks[c];
  }
  if (n_a > 0)
    if (int rc = launch_append(ao, n_a, frame_idx, 0.f, st)) return rc;
  if (n_c > 0)
    if (int rc = launch_clamp(cb, n_c, st, commit)) return rc;
  return VFN_OK;
}

int vfn_bank_update_finish(vfn_bank* banks, int32_t obj_n, vfn_update_io* io, const int32_t* h_pinned) {
  VFN_CHECK_ARG(banks && io && h_pinned && obj_n >= 1 && obj_n <= 4, "bank_update_finish: bad args");
  for (int c = 0; c < obj_n; ++c) {
    if (!io[c].deferred) continue;
    io[c].n_merge = h_pinned[4 * c + 0];
    io[c].n_runs = h_pinned[4 * c + 1];
    io[c].n_append = h_pinned[4 * c + 2];
    if (banks[c].n_live) {
      // device-resident count: the plan kernel staged the exact live count after this update's append
      const int64_t n_next = h_pinned[4 * c + 3];
      VFN_CHECK_ARG(io[c].n_append >= 0 && n_next >= banks[c].n_min && n_next <= banks[c].n + io[c].n_append &&
                        n_next <= banks[c].cap,
                    "bank_update_finish: counts of object %d are not valid (was the event waited on?)", c);
      banks[c].n = banks[c].n_min = n_next;
    } else {
      VFN_CHECK_ARG(io[c].n_append >= 0 && banks[c].n + io[c].n_append <= banks[c].cap,
                    "bank_update_finish: counts of object %d are not valid (was the event waited on?)", c);
      banks[c].n += io[c].n_append;
    }
    io[c].deferred = 0;
    // algorithmic bytes of the deferred append (the launch could only account an upper bound: it recorded 0)
    vfn_profile_add_work(PROF_APPEND, 2.0 * 4.0 * (banks[c].d_key + banks[c].d_val + 2) * (double)io[c].n_append);
  }
  return VFN_OK;
}

}  // extern "C"
