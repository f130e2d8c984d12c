/* bm_oracle_int.h -- TEST INFRASTRUCTURE ONLY: helpers shared by the oracle's translation units (byte reader of a BLOB). */
#ifndef BM_ORACLE_INT_H_INCLUDED
#define BM_ORACLE_INT_H_INCLUDED
#include "bm_oracle.h"

/* bm::decoder (src/encoding.h): little-endian byte reader; reading past the end parks p beyond end (callers test p > end) */
typedef struct { const uint8_t* p; const uint8_t* end; } rd_t;
static inline uint32_t rd8(rd_t* r)  { if (r->p + 1 > r->end) { r->p = r->end + 1; return 0; } return *r->p++; }
static inline uint32_t rd16(rd_t* r) { uint32_t a = rd8(r); return a | (rd8(r) << 8); }
static inline uint32_t rd32(rd_t* r) { uint32_t a = rd16(r); return a | (rd16(r) << 16); }
static inline uint64_t rd64(rd_t* r) { uint64_t a = rd32(r); return a | ((uint64_t)rd32(r) << 32); }

/* bm_oracle_entropy.c */
int orc_entropy_token(rd_t* r, uint32_t bt, uint32_t* tb, int* is_gap);
int orc_sblock_token(rd_t* r, uint32_t bt, uint32_t* arr, uint32_t* len_out, uint32_t* sb);
#endif
