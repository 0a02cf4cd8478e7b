/*
 * vfn.h - C ABI of libvfn_sm100a.so: the B200 (sm_100a) implementation of V-FloodNet's AFB-URR
 * memory-propagation hot path (feature-bank read, bank update, uncertain-region refinement).
 *
 * The reference (xmlyqing00/V-FloodNet) has no FFI layer: its "operator API" for this path is three
 * Python objects (FeatureBank, Matcher, Decoder).  Each entry point below names the reference lines it
 * replaces; vfloodnet_b200/*.py re-creates those Python objects on top of this ABI via ctypes.
 *
 * Conventions
 *   - every pointer named d_* / inside vfn_bank is a DEVICE pointer; h_* is a HOST pointer (pinned preferred)
 *   - `stream` is a cudaStream_t passed as void*
 *   - return value: 0 = ok, <0 = error (VFN_E_*); vfn_last_error() returns a thread-local message
 *   - no allocation, no host synchronisation and no exceptions inside the library unless stated;
 *     scratch memory is passed in (`ws`, sized by the matching *_workspace_bytes query)
 *   - layouts: "dm" = dimension-major (d, n) as the reference stores keys/values (FeatureBank.py:29-30),
 *              "em" = entry-major (n, d): one bank slot / one query per contiguous row (the HBM layout)
 */
#ifndef VFN_H_
#define VFN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VFN_VERSION 102

#define VFN_OK 0
#define VFN_E_ARG (-1)       /* bad argument / unsupported shape */
#define VFN_E_CAPACITY (-2)  /* bank capacity or workspace too small */
#define VFN_E_CUDA (-3)      /* CUDA runtime error, see vfn_last_error() */
#define VFN_E_UNSUPPORTED (-4)

/* One object's feature bank in HBM, entry-major.  Owned by the caller (torch allocations in the Python host). */
typedef struct vfn_bank {
  int32_t d_key;   /* 128 in AFB_URR (AFB_URR.py:250) */
  int32_t d_val;   /* 512 */
  int64_t cap;     /* allocated slots */
  int64_t n;       /* live slots (host-tracked) */
  float* keys;     /* (cap, d_key)  raw keys      == reference fb.keys[i].T   */
  float* values;   /* (cap, d_val)  raw values    == reference fb.values[i].T */
  float* info;     /* (cap, 2)      [frame added, sum log(cnt+1)]  (FeatureBank.py:34-35) */
  float* nk;       /* (cap, d_key)  cached NF.normalize(keys, dim=0) (FeatureBank.py:63), exact fp32: SIMT match + re-score */
  uint16_t* nkh;   /* (cap, d_key)  fp16 hi of 16*nk  - tensor-core match operand, NULL if unused */
  uint16_t* nkl;   /* (cap, d_key)  fp16 lo  (16*nk - hi) */
  uint16_t* kh;    /* (cap, d_key)  fp16 hi of keys   - tensor-core read operand, NULL if unused */
  uint16_t* kl;    /* (cap, d_key)  fp16 lo  (key - hi) */
  uint16_t* vh;    /* (cap, d_val)  fp16 hi of values */
  uint8_t* v8;     /* (cap, d_val)  e4m3(value)        - partner of the P-residual pass */
  uint8_t* vl;     /* (cap, d_val)  e5m2(value - vh)   - partner of the e4m3(P) pass */
  int32_t* cnt;    /* (cap)         usage-count scratch, all zero between reads */
  /* Device-resident live count (optional; NULL = `n` is exact and the kernels use it).  n_live[0] = live slots,
   * n_live[1] = count staged by the update in flight.  When non-NULL the kernels take the live count from n_live[0]
   * and the host fields become bounds: n_min <= n_live[0] <= n.  `n` sizes grids and workspaces, `n_min` bounds the
   * work partition.  This lets a caller queue read(t+1) behind update(t) without reading |A| back first
   * (the reference's nonzero() sync, FeatureBank.py:100).  tcgen05 paths only (d_key=128, d_val=512, impl != 1).
   * vfn_bank_update keeps n_live current on every path; after vfn_bank_append_rows / vfn_bank_compact called directly,
   * call vfn_bank_set_live. */
  int32_t* n_live;
  int64_t n_min;
} vfn_bank;

int vfn_version(void);
const char* vfn_last_error(void);
/* 1 if the running device is compute capability 10.x (tcgen05/TMEM/TMA kernels usable) */
int vfn_device_is_sm100(void);
/* sizeof(vfn_bank) / sizeof(vfn_update_io) as compiled into the library: a binding checks them against its own layout
 * before the first call (a stale library with other struct layouts must not be driven silently) */
int vfn_abi_sizeof_bank(void);
int vfn_abi_sizeof_update_io(void);

/* ---- candidate / query preparation ------------------------------------------------------------
 * (d, n) dm  ->  (n, d) em copies: raw and L2-normalised (x / max(|x|, 1e-12)), + optional fp16 hi/lo of raw*scale
 * (of normalised*scale when d_normed_em is non-NULL: the tensor-core match operand).
 * Replaces NF.normalize(prev_key, dim=0) / NF.normalize(prev_value, dim=0) (FeatureBank.py:64,88) and the
 * transposes implied by keys[i].transpose(0,1) (AFB_URR.py:144).  Any output pointer may be NULL. */
int vfn_prep_rows(const float* d_src_dm, int32_t d, int64_t n, float* d_raw_em, float* d_normed_em,
                  uint16_t* d_hi_em, uint16_t* d_lo_em, float scale, void* stream);

/* ---- bank ingest: init_bank / append / update's append step ------------------------------------
 * Copies rows sel[0..n_sel) (ascending candidate positions; sel==NULL => 0..n_sel-1) of the em candidate
 * arrays into slots [bank->n, bank->n+n_sel), writes info rows (info0, info1) and all derived arrays.
 * d_n_sel (device int, may be NULL) overrides n_sel_upper with a device-side count (no host sync).
 * Replaces FeatureBank.init_bank (FeatureBank.py:27-36), append (:38-51), and update's cat (:105-111). */
int vfn_bank_append_rows(const vfn_bank* bank, const float* d_ck_em, const float* d_cv_em, const float* d_nck_em,
                         const int32_t* d_sel, int64_t n_sel_upper, const int32_t* d_n_sel, float info0, float info1,
                         void* stream);

/* n_live[0] = n_live[1] = n (one tiny kernel on `stream`); no-op when bank->n_live is NULL */
int vfn_bank_set_live(const vfn_bank* bank, int64_t n, void* stream);

/* Recompute derived arrays (nk, nkh/nkl, kh/kl, vh/v8/vl) of slots [first, first+count) from keys/values. */
int vfn_bank_refresh(const vfn_bank* bank, int64_t first, int64_t count, void* stream);

/* ---- memory read: Matcher.forward (AFB_URR.py:136-178) ------------------------------------------
 * out: (obj_n, 2*d_val, hw) fp32 = [softmax_i(K_i.q_j/sqrt(d_key)) readout ; q_out] per object
 * (== reference (1,obj_n,1024,HW), bs=1).  If update_bank: info[:,1] += log(cnt+1) with
 * cnt_i = #{j : p_ij > thres_valid} (AFB_URR.py:161-174).  d_lse (obj_n, hw) optional: natural-log LSE.
 * impl: bits 0-7: 0 = auto (tcgen05 when shapes allow), 1 = fp32 SIMT kernels, 2 = tcgen05 kernels;
 *       bit 8 (VFN_Q_IN_EM): d_q_in_dm points to an ENTRY-MAJOR (hw, d_key) query tensor - what vfn_keyvalue writes -
 *       instead of the reference's (d_key, hw); the transpose of the query preparation is skipped (same values). */
#define VFN_Q_IN_EM 0x100
size_t vfn_memread_workspace_bytes(int32_t obj_n, int64_t n_max, int64_t hw, int32_t d_key, int32_t d_val);
int vfn_memread(const vfn_bank* banks, int32_t obj_n, const float* d_q_in_dm, const float* d_q_out_dm, int64_t hw,
                float thres_valid, int32_t update_bank, float* d_out, float* d_lse, void* d_ws, size_t ws_bytes,
                int32_t impl, void* stream);

/* Split-memory form for a bank sharded over GPUs: phase A leaves per-query (max, sum exp(s-max)) of the
 * LOCAL slots in d_ml (obj_n, hw, 2) [natural-log domain, s = <k,q>/sqrt(d_key)]; the caller stacks the partial sets of
 * all ranks and reduces them with vfn_lse_combine into the global LSE d_lse (obj_n, hw); phase B (same workspace,
 * called after phase A) computes the local partial readout (obj_n, d_val, hw) + usage counts against that global LSE.
 * Partial readouts are summed across ranks by the caller (all-reduce). */
int vfn_memread_phase_a(const vfn_bank* banks, int32_t obj_n, const float* d_q_in_dm, int64_t hw, float* d_ml,
                        void* d_ws, size_t ws_bytes, int32_t impl, void* stream);
int vfn_memread_phase_b(const vfn_bank* banks, int32_t obj_n, const float* d_q_in_dm, int64_t hw, const float* d_lse,
                        float thres_valid, int32_t update_bank, float* d_partial_out, void* d_ws, size_t ws_bytes,
                        int32_t impl, void* stream);
/* d_ml: n_parts stacked partial sets (n_parts, n_rows, 2) -> d_lse (n_rows) = M + log(sum_s l_s exp(m_s - M)) */
int vfn_lse_combine(const float* d_ml, int32_t n_parts, int64_t n_rows, float* d_lse, void* stream);

/* ---- exchange steps of a bank sharded over the GPUs of one node, over PEER memory ---------------------
 * h_peer_*: HOST array of n_parts DEVICE pointers, entry r = rank r's buffer as mapped into THIS process (CUDA IPC /
 * symmetric memory; entry `self` is the local buffer).  The kernels read the peers' buffers straight over NVLink - there
 * is no library collective on the data path.  The caller brackets each call with a cross-GPU barrier on `stream`
 * (all producers done before, all consumers done after the buffers are overwritten).  n_parts <= 16.
 *   lse_combine_peers:   rank r's buffer = its local (m, l) pairs (n_rows, 2) from vfn_memread_phase_a; d_lse (n_rows) =
 *                        the global log-sum-exp (same arithmetic and order on every rank)
 *   reduce_peers:        d_out[first .. first+count) = sum_r peer_r[first .. first+count), summed in rank order (fp32,
 *                        deterministic: every rank that reduces the same range gets the same bits); first, count % 4 == 0
 *   gather_peers:        d_out[i] = peer_r[i] for i in slice r (slice floats per rank, the last rank takes the rest):
 *                        the all-gather half of a two-shot reduction (each rank reduced its own slice in place with
 *                        vfn_reduce_peers; `self` is only range-checked)
 *   match_pack / match_combine_peers: the arg-max combine of the cosine match (FeatureBank.py:66-68) across shards.
 *                        pack: pair[q] = {corr bits, global sequence id of the matched local slot} (16 bytes per query);
 *                        combine: best corr and, among the ranks reaching it, the LOWEST sequence id (= the reference's
 *                        lowest index: shards keep insertion order); a NaN score makes (NaN, INT64_MAX). */
int vfn_lse_combine_peers(const void* const* h_peer_ml, int32_t n_parts, int64_t n_rows, float* d_lse, void* stream);
int vfn_reduce_peers(const void* const* h_peer_part, int32_t n_parts, int64_t first, int64_t count, float* d_out,
                     void* stream);
int vfn_gather_peers(const void* const* h_peer_buf, int32_t n_parts, int32_t self, int64_t slice, int64_t total,
                     float* d_out, void* stream);
int vfn_match_pack(const float* d_corr, const int32_t* d_idx, const int64_t* d_seq_of_slot, int64_t n_local, int64_t hw,
                   void* d_pair, void* stream);
int vfn_match_combine_peers(const void* const* h_peer_pair, int32_t n_parts, int64_t hw, float* d_best_corr,
                            int64_t* d_best_seq, void* stream);

/* ---- bank update: FeatureBank.update (FeatureBank.py:53-115) -----------------------------------
 * match: j*_q = argmax_i <nk_i, nck_q>, ties -> lowest i; c*_q = that maximum.  (FeatureBank.py:63-68) */
size_t vfn_bank_match_workspace_bytes(int64_t n, int64_t hw);
int vfn_bank_match(const vfn_bank* bank, const float* d_nck_em, int64_t hw, int32_t* d_match_idx, float* d_match_corr,
                   void* d_ws, size_t ws_bytes, int32_t impl, void* stream);

/* plan: from (j*, c*) build   merge set S = {q : c*>thres} as (slot,q) pairs sorted by (slot, q),
 * run offsets of equal-slot runs (= unique touched slots, ascending), append set A = {q : c*<=thres} ascending.
 * d_counts[0..3] = {n_merge, n_runs, n_append, n_next}; also copied to h_counts (pinned) if non-NULL.
 * n_next = live count after the append = n_live[0] + n_append when the plan is run for a bank with a device-resident
 * count (vfn_bank_update), else 0.
 * Replaces nonzero / unique / index plumbing of FeatureBank.py:71-73,100. */
size_t vfn_bank_plan_workspace_bytes(int64_t hw);
int vfn_bank_plan(const int32_t* d_match_idx, const float* d_match_corr, int64_t hw, float thres_close,
                  int32_t* d_merge_q, int32_t* d_merge_slot, int32_t* d_run_off, int32_t* d_append_q,
                  int32_t* d_counts, int32_t* h_counts, void* d_ws, size_t ws_bytes, void* stream);

/* merge: for each run (slot u, members q ascending): mean of normalised candidates (scatter_mean, sequential
 * ascending-q summation), then  row_u <- |row_u| * ((1-r)*row_u/|row_u| + r*mean)  for keys and values
 * (FeatureBank.py:76-97), and refresh of the derived arrays of slot u. */
int vfn_bank_merge(const vfn_bank* bank, const float* d_nck_em, const float* d_ncv_em, const int32_t* d_merge_q,
                   const int32_t* d_merge_slot, const int32_t* d_run_off, const int32_t* d_counts, int64_t hw,
                   float update_rate, void* stream);

/* evict plan: FeatureBank.remove's threshold search (FeatureBank.py:117-138), entirely on device.
 * LFU_i = info[i,1]/(frame_idx - info[i,0]); T = int(min LFU)+1; keep LFU > T; while
 * class_budget - kept - request_n < 0: T = int(min of survivors)+1.
 * d_plan (int32[72]) = {status, kept, n_iter, T_final, thresholds[0..63]...}; mirrored to h_plan if non-NULL.
 * The search runs until the balance is met (no iteration cap, like the reference); n_iter counts every threshold tried,
 * of which the first 64 are recorded.
 * status: 0 ok, 1 = survivors empty while balance<0 (reference raises), 2 = non-finite LFU minimum (reference raises). */
int vfn_bank_evict_plan(const vfn_bank* bank, float frame_idx, double class_budget, int64_t request_n,
                        int32_t* d_plan, int32_t* h_plan, float* d_lfu_scratch, void* stream);
/* compaction: order-preserving copy of the slots with LFU > T_final (read from d_plan) from src into dst
 * (all arrays of the bank).  dst->cap >= kept.  ws: vfn_bank_compact_workspace_bytes(src->n). (FeatureBank.py:127-131) */
size_t vfn_bank_compact_workspace_bytes(int64_t n);
int vfn_bank_compact(const vfn_bank* src, const vfn_bank* dst, const float* d_lfu, const int32_t* d_plan, void* d_ws,
                     size_t ws_bytes, void* stream);

/* info[:,1] = clamp(info[:,1], 0, 1e5) over slots [0, n)  (FeatureBank.py:115) */
int vfn_bank_clamp_info(const vfn_bank* bank, int64_t n, void* stream);

/* ---- whole update in one call: FeatureBank.update (FeatureBank.py:53-115) for all objects ----------
 * Runs, for every object: candidate preparation, match, plan, merge, [LFU eviction if class_budget < n + |A|:
 * evict plan + compaction into alts[c]], append, clamp - the exact sequence of the fine-grained entry points above,
 * issued from C++ so that a frame costs one host call instead of ~25.  This entry point DOES synchronise the stream
 * (once for the append counts, once more if any object evicts): those are the reference's own implicit syncs
 * (`nonzero`, `int(LFU.min())`, FeatureBank.py:71,100,123).
 * banks[c].n is updated in place.  If object c evicted, its live data moved into alts[c]: the two structs are
 * swapped and io[c].swapped = 1 so the caller swaps its buffer ownership.  Capacity contract: banks[c].cap >= n + hw,
 * alts[c].cap >= n (alts[c].keys may be NULL when the caller knows no eviction can happen: then an eviction is an error).
 * io[c].d_* are caller-owned device buffers of hw (run_off: hw + 1) elements holding the decisions of this update. */
typedef struct vfn_update_io {
  const float* d_prev_key_dm;    /* (d_key, hw) candidate keys    (memorize() output, AFB_URR.py:255-272) */
  const float* d_prev_value_dm;  /* (d_val, hw) candidate values; both are (hw, d) when prev_layout == 1 */
  int32_t* d_match_idx;          /* (hw) j*            */
  float* d_match_corr;           /* (hw) c*            */
  int32_t* d_merge_q;            /* (hw) merge pairs sorted by (slot, q) */
  int32_t* d_merge_slot;         /* (hw) */
  int32_t* d_run_off;            /* (hw + 1) */
  int32_t* d_append_q;           /* (hw) append set, ascending */
  /* host results */
  int32_t n_merge, n_runs, n_append;
  int32_t evicted;               /* 1 if remove() ran for this object */
  int32_t swapped;               /* 1 if banks[c] / alts[c] were exchanged */
  int32_t evict_status;          /* 0 ok, 1 = survivors empty (reference raises), 2 = non-finite LFU minimum */
  int32_t kept, n_iter;
  int32_t thresholds[64];        /* T sequence of remove(): the first min(n_iter, 64) thresholds */
  int64_t n_before;              /* bank size before this update */
  int32_t deferred;              /* 1: counts not read back yet - call vfn_bank_update_finish() after the event */
  int32_t prev_layout;           /* input: 0 = d_prev_* are (d, hw) dimension-major (the reference's memorize() output),
                                  * 1 = (hw, d) entry-major (vfn_keyvalue's *_em outputs: no transpose in the preparation) */
} vfn_update_io;

size_t vfn_bank_update_workspace_bytes(int32_t obj_n, int64_t n_max, int64_t hw, int32_t d_key, int32_t d_val);
/* h_pinned: obj_n * 80 int32 of pinned host memory (counts + eviction plans land there).
 * defer_event: NULL, or a cudaEvent_t.  With an event, an update in which no object can reach its budget
 * (n + hw <= class_budget for all objects, so FeatureBank.py:102 cannot fire) does NOT synchronise the stream: the
 * event is recorded once the counts are on their way to h_pinned, io[c].deferred is set, banks[c].n is left unchanged,
 * and the caller completes the update with vfn_bank_update_finish() after cudaEventSynchronize(defer_event), before
 * anything else touches the banks or h_pinned.  Updates that may evict keep the reference's synchronous behaviour.
 * Banks with a device-resident count (n_live) may be passed with n = upper bound / n_min = lower bound: the deferral
 * test uses the upper bound, every kernel of the update reads the live count on the device, and the caller may issue
 * the next read and the next update (bounds advanced by hw) without finishing this one. */
int vfn_bank_update(vfn_bank* banks, vfn_bank* alts, int32_t obj_n, vfn_update_io* io, int64_t hw, float frame_idx,
                    float update_rate, float thres_close, double class_budget, void* d_ws, size_t ws_bytes,
                    int32_t* h_pinned, int32_t impl, void* defer_event, void* stream);
/* host-only: reads |merge|, |runs|, |append| of a deferred update from h_pinned, fills io and advances banks[c].n
 * (banks with n_live: sets n = n_min = the exact live count the plan kernel staged, so any number of deferred updates
 * may be in flight and only the last one is finished). */
int vfn_bank_update_finish(vfn_bank* banks, int32_t obj_n, vfn_update_io* io, const int32_t* h_pinned);

/* ---- KeyValue head: KeyValue.forward (AFB_URR.py:94-111), SURVEY 8(f) n3 ---------------------------------------
 * Key = conv3x3(x, wk (d_key, c_in, 3, 3)) + bk, Value = conv3x3(x, wv (d_val, c_in, 3, 3)) + bv, stride 1, zero
 * padding 1, on x (B, c_in, h, w) fp32 (r4 of EncoderQ / EncoderM), as one tcgen05 implicit GEMM.
 * pack_weights (once per model; again after the weights change): fp16 hi/lo operand arrays [tap][n][c] + bias into
 *   d_packed (vfn_kv_packed_weights_bytes); bk / bv may be NULL.
 * keyvalue: any of the four outputs may be NULL.  *_em (B, h*w, d) entry-major = the layout of the read's query
 *   operand (vfn_memread with VFN_Q_IN_EM) and of the update's candidate rows (vfn_update_io.prev_layout = 1);
 *   *_dm (B, d, h*w) = the reference's `key.view(*key.shape[:2], -1)` (AFB_URR.py:105-109).
 *   passes: 3 = fp16 hi/lo splits of both operands, hi*hi + lo*hi + hi*lo, fp32 accumulate (the result of a true-fp32
 *   convolution up to summation order); 1 = hi*hi only (11-bit operands: the class of cuDNN's TF32 default).
 * Shapes: c_in % 64 == 0, d_key % 32 == 0, d_val % 32 == 0, (d_key + d_val) % 320 == 0 (128 + 512 of AFB_URR.py:250). */
size_t vfn_kv_packed_weights_bytes(int32_t c_in, int32_t d_key, int32_t d_val);
int vfn_kv_pack_weights(const float* d_wk, const float* d_bk, const float* d_wv, const float* d_bv, int32_t c_in,
                        int32_t d_key, int32_t d_val, void* d_packed, void* stream);
size_t vfn_keyvalue_workspace_bytes(int32_t B, int32_t c_in, int32_t h, int32_t w, int32_t d_key, int32_t d_val);
int vfn_keyvalue(const float* d_x, int32_t B, int32_t c_in, int32_t h, int32_t w, const void* d_packed, int32_t d_key,
                 int32_t d_val, int32_t passes, float* d_key_em, float* d_val_em, float* d_key_dm, float* d_val_dm,
                 void* d_ws, size_t ws_bytes, void* stream);

/* ---- URR: non-convolution parts of Decoder.forward (AFB_URR.py:214-237, myutils/data.py:42-48) -----
 * pre:  p (obj_n,2,h/2,w/2) coarse logits from pred2, r1 (obj_n or 1, c, h, w)  ->
 *       p_up (obj_n,2,h,w) bilinear x2 of p; seg (obj_n,h,w) object-normalised foreground prob (rough_seg);
 *       unc (h,w) = exp(1 - top1/(top2+1e-8)) over objects (identical for every object, bs=1);
 *       conf (obj_n,h,w) 7x7 max of seg; avg (obj_n,h,w) 7x7 mean of seg (zero pad, /49);
 *       local_match (obj_n,2c,h,w) = [r1 ; box7(r1*seg)/49 / (avg + 1e-8)]   (input of local_convFM)
 * post: prob (obj_n,2h,2w) = softmax(bilinear_x2(p_up + unc * (conf * q_local)))[:,1]
 * r1_obj_stride: elements between objects in r1 (0 => the same r1 for every object = the reference's expand). */
int vfn_urr_pre(const float* d_p, const float* d_r1, int64_t r1_obj_stride, int32_t obj_n, int32_t c, int32_t h,
                int32_t w, float* d_p_up, float* d_seg, float* d_unc, float* d_conf, float* d_avg,
                float* d_local_match, void* stream);
int vfn_urr_post(const float* d_p_up, const float* d_unc, const float* d_conf, const float* d_q_local, int32_t obj_n,
                 int32_t h, int32_t w, float* d_prob, void* stream);

/* ---- frame-loop tail: what follows fb.update in the reference's loop, kept on the device ---------
 * resize_argmax: pred (H,W) u8 = argmax_c bicubic(pred_mask (obj_n,h,w) -> (H,W)); ties to the lowest class
 *   replaces `TF.resize(pred_mask, ori_size, BICUBIC)` + `torch.argmax(pred[0], dim=0).cpu()` (test_video_seg.py:114-115).
 *   antialias != 0: F.interpolate(mode='bicubic', antialias=True) (a = -0.5, truncated + re-normalised window;
 *   torchvision >= 0.17 TF.resize on tensors); 0: a = -0.75, clamped indices (torchvision 0.9.1, README pin).
 * largest_component: mask (H,W) u8 = the largest 8-connected component of pred != 0; equal sizes resolve like cv2's
 *   CCL_GRANA label order (first 2x2 block in raster order); an empty pred gives an all-ones mask.  Replaces
 *   `myutils.postprocessing_pred` (myutils/data.py:19-39) for binary predictions (obj_n == 2).
 *   d_stats int32[4] = {foreground pixels, components, size of the kept one, its block-raster root id or -1}.
 * waterlevel: for key point t = (x, y) = d_key_pts[2t], d_key_pts[2t+1]: first row y' > y with mask[y'][x] ==
 *   water_label_id -> d_level[t] = y' - y (NaN when that is 1); no such row, or x outside the image: d_level[t] is left
 *   as it is (the reference carries the previous frame's estimate forward).  estimation/reference_tracking.py:190-204.
 * vfn_frame_tail: the three in sequence on one stream (7 launches, no host synchronisation). */
size_t vfn_tail_workspace_bytes(int32_t H, int32_t W);
int vfn_tail_resize_argmax(const float* d_pred_mask, int32_t obj_n, int32_t h, int32_t w, int32_t H, int32_t W,
                           int32_t antialias, uint8_t* d_pred, void* stream);
int vfn_tail_largest_component(const uint8_t* d_pred, int32_t H, int32_t W, uint8_t* d_mask, int32_t* d_stats,
                               void* d_ws, size_t ws_bytes, void* stream);
int vfn_tail_waterlevel(const uint8_t* d_mask, int32_t H, int32_t W, const int32_t* d_key_pts, int32_t n_pts,
                        int32_t water_label_id, float* d_level, void* stream);
int vfn_frame_tail(const float* d_pred_mask, int32_t obj_n, int32_t h, int32_t w, int32_t H, int32_t W,
                   int32_t antialias, const int32_t* d_key_pts, int32_t n_pts, int32_t water_label_id,
                   uint8_t* d_pred, uint8_t* d_mask, int32_t* d_stats, float* d_level, void* d_ws, size_t ws_bytes,
                   void* stream);

/* ---- measurement hooks (bench.py) ----------------------------------------------------------------
 * While enabled, the library brackets its dominant kernels with CUDA events recorded on the launching stream.
 * kinds: 0 read phase A, 1 read phase B, 2 match, 3 compaction move, 4 merge, 5 append, 6 URR local.
 * collect: h_out[3*k + {0,1,2}] = {launches, total ms, total algorithmic work (flop or bytes)} (host sync on the events).
 * vfn_launch_count(): number of kernels this library has launched in this process. */
int vfn_profile_enable(int32_t on);
int vfn_profile_collect(double* h_out, int32_t n_kinds);
int vfn_profile_add_work(int32_t kind, double work);
int64_t vfn_launch_count(void);

/* ---- self-test hooks for the tcgen05/TMA building blocks (used by tests/, not by the product path) ---- */
/* d_ptr != NULL (74 x 64 x 8 int64, zeroed): the CTA-pair phase-B kernel stamps clock64 per cluster and work item:
 * {item start, P of the first tile ready, P of the last tile ready, item end, tiles} (tests/debug_item_times.py) */
int vfn_debug_set_tstamp(long long* d_ptr);
/* d_ptr != NULL: the next tcgen05 read launches dump the first S^T tile of CTA 0 (128 x tile floats) there. */
int vfn_debug_set_dump(float* d_ptr);
/* bit mask, default 3: bit 0 = CTA-pair (cta_group::2) phase B, bit 1 = CTA-pair score scan (phase A, match);
 * 0 selects the single-CTA kernels (cross-check in tests/) */
int vfn_debug_set_pair(int32_t mask);
/* streaming (warp-shuffle, register-ring) URR local kernel when w % 4 == 0: 1 = two objects per warp, 2 = one object per
 * warp; 0: tiled shared-memory kernel.  All three are bit-identical (tests/). */
int vfn_debug_set_urr_stream(int32_t on);

/* 1 (default): kernels that support it are queued with programmatic dependent launch (their CTAs are placed while the
 * predecessor drains); 0: plain stream-ordered launches */
int vfn_debug_set_pdl(int32_t on);
/* tail labelling cross-check: bit 1 = no per-CTA size aggregation (every warp adds to global memory) */
int vfn_debug_set_tail(int32_t flags);

#ifdef __cplusplus
}
#endif
#endif /* VFN_H_ */
