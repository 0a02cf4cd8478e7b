// Interface between the dispatch code (vfn_simt.cu) and the tcgen05/TMEM/TMA kernels (vfn_tc.cu).
#pragma once
#include "vfn_common.cuh"

namespace vfn {

// true when the tensor-core read kernels support these dims (d_key = 128, d_val = 512: AFB_URR.py:250)
bool tc_shapes_ok(int d_key, int d_val);
void tc_pick_splits(int obj_n, int64_t n_max, int64_t hw, int* split_a, int* split_b);
size_t tc_workspace_bytes(int obj_n, int64_t hw);
// phase A: (m, l) partials [natural-log domain] for every (object, split, query): part[(obj*split_a + s)*hw + j]
// split_a / split_b are the workspace bounds; the number of pieces actually used comes back in *pieces_out and is the
// n_split the combine kernels must use - unless the banks carry device-resident live counts (vfn_bank::n_live): then
// the split is chosen on the device from the live sizes and *pieces_dev_out points to it (NULL otherwise).
// q_em != 0: q_in is (hw, d_key) entry-major (vfn_keyvalue's output) instead of (d_key, hw)
int tc_phase_a(const vfn_bank* banks, int obj_n, const float* q_in_dm, int64_t hw, int split_a, float2* part,
               char* ws_tc, cudaStream_t st, int* pieces_out, const int32_t** pieces_dev_out, int q_em = 0);
// phase B: partial readouts po[((obj*split_b + s)*d_val + c)*hw + j] and usage counts into bank.cnt
int tc_phase_b(const vfn_bank* banks, int obj_n, const float* q_in_dm, int q_em, int64_t hw, int split_b,
               const float* lse, float thres_valid, int update_bank, float* po, char* ws_tc, cudaStream_t st,
               int* pieces_out, const int32_t** pieces_dev_out);

// cosine match for obj_n banks in one launch: fp16x3 tcgen05 scores -> per-(piece, query, column group) near-tie
// candidates in the workspace -> exact fp32 re-score (same FMA chain as the SIMT kernel) -> idx_out[o] / corr_out[o].
// ws: tc_match_workspace_bytes(obj_n, hw).  cand_split != 0: the fp16 hi/lo of 16 * normalised candidates were already
// written to tc_match_cand_hi/lo(ws, ...) (zero padded to a multiple of 128 rows) by the preparation kernel.
size_t tc_match_workspace_bytes(int obj_n, int64_t hw);
// rows of an A-operand array (fp16 hi or lo, 128 columns): hw rounded up to a whole number of query-tile pairs; the rows
// beyond hw must be zero (PrepJob::n_pad makes the preparation launch write them)
int64_t tc_operand_rows(int64_t hw);
uint16_t* tc_match_cand_hi(char* ws, int obj_n, int64_t hw, int obj);
uint16_t* tc_match_cand_lo(char* ws, int obj_n, int64_t hw, int obj);
int tc_match(const vfn_bank* banks, int obj_n, const float* const* nck_em, int64_t hw, char* ws, int cand_split,
             int32_t* const* idx_out, float* const* corr_out, cudaStream_t st);

}  // namespace vfn
