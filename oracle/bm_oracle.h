/*
 * bm_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference (tlk00/BitMagic v9.2.1) algorithms for the
 * block-level set-algebra hot path and the rank/select path.  It exists to CHECK the
 * CUDA path; nothing in bitmagic_b200/ may call, link or import it.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Parity pinning: tests/test_oracle_vs_reference.py checks every function here against
 * the unmodified reference built from /root/reference/src (oracle/_ref/libbmref.so,
 * recipe in oracle/Makefile) and against the committed fixtures in tests/golden/
 * (generated from that reference build by tests/golden/make_golden.py).
 */
#ifndef BM_ORACLE_H_INCLUDED
#define BM_ORACLE_H_INCLUDED
#include <stdint.h>
#include "../include/bmb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- single block primitives ---- */
uint32_t orc_bit_block_count(const uint32_t* blk);
uint64_t orc_block_digest(const uint32_t* blk);
uint32_t orc_bit_block_calc_change(const uint32_t* blk);
uint32_t orc_bit_to_gap(uint16_t* dest, const uint32_t* blk);
void     orc_gap_convert_to_bitset(uint32_t* blk, const uint16_t* gap);
void     orc_gap_add_to_bitset(uint32_t* blk, const uint16_t* gap);
void     orc_gap_and_to_bitset(uint32_t* blk, const uint16_t* gap);
void     orc_gap_sub_to_bitset(uint32_t* blk, const uint16_t* gap);
void     orc_gap_xor_to_bitset(uint32_t* blk, const uint16_t* gap);
uint32_t orc_gap_bit_count(const uint16_t* gap);
uint32_t orc_gap_bfind(const uint16_t* gap, uint32_t pos, uint32_t* is_set);
uint32_t orc_gap_bit_count_range(const uint16_t* gap, uint32_t left, uint32_t right);
uint32_t orc_bit_block_count_range(const uint32_t* blk, uint32_t left, uint32_t right);

/* ---- whole-set aggregation over a packed (column-major) set ----
 * outputs are per column c = nb - nb_from (any pointer may be NULL):
 *   kind[c], popcnt[c], digest[c], nruns[c], blocks[c*2048..] (logical bits of the result,
 *   always filled, zeros / ones for NULL / FULL), gaps[c*1280..] (GAP form when kind == GAP) */
int orc_aggregate(const bmb200_packed_set* set, const bmb200_agg_args* args,
                  uint8_t* kind, uint32_t* popcnt, uint64_t* digest, uint32_t* nruns,
                  uint32_t* blocks, uint16_t* gaps);

/* ---- bm::deserialize of a serialized bvector BLOB into an empty vector (explicit-length token subset) ----
 * kind[n_cols]; blocks[n_cols*2048] logical bits (may be NULL); gaps[n_cols*1280] GAP form of the GAP-kind columns (may be NULL) */
int orc_deserialize(const uint8_t* blob, uint64_t size, uint32_t n_cols, uint8_t* kind, uint32_t* blocks, uint16_t* gaps);
/* optional: 256 counters, one per token type met by orc_deserialize (index 0x80 = packed short zero run); NULL switches it off */
void orc_set_token_hist(uint32_t* hist);

/* ---- sparse-vector scanner searches (bmb200_scan): n_values * n_cols columns, value-major ---- */
int orc_scan(const bmb200_packed_set* set, const bmb200_scan_args* args,
             uint8_t* kind, uint32_t* popcnt, uint64_t* digest, uint32_t* nruns, uint32_t* blocks, uint16_t* gaps);

/* ---- synthetic benchmark inputs on the host (bm_synth.c): the same counter-based generator as bmb200_synth_set ---- */
typedef struct orc_synth orc_synth;
int  orc_synth_create(uint32_t n_vec, uint32_t n_blocks, const double* density, const uint64_t* seed, int optimize, int threads,
                      orc_synth** out);
void orc_synth_packed(const orc_synth* s, bmb200_packed_set* out);      /* view into s; valid until orc_synth_free */
void orc_synth_free(orc_synth* s);
void orc_synth_force_scalar(int on);                                    /* 1 = never take the AVX-512 form of the generator */

/* ---- rank / select over vector `vec` of a packed set ---- */
int orc_rs_build(const bmb200_packed_set* set, uint32_t vec,
                 uint32_t* bcount, uint64_t* sub_count, uint64_t* sb_count);
int orc_rank_batch(const bmb200_packed_set* set, uint32_t vec,
                   const uint64_t* pos, uint64_t n, uint64_t* out);
int orc_select_batch(const bmb200_packed_set* set, uint32_t vec,
                     const uint64_t* rank, uint64_t n, uint64_t* pos, uint8_t* found);

/* expand block (vec, nb) of the set to its logical 2048 words */
void orc_expand_block(const bmb200_packed_set* set, uint32_t vec, uint32_t nb, uint32_t* out);

#ifdef __cplusplus
}
#endif
#endif
