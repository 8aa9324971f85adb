/* sw_oracle.c -- see sw_oracle.h. TEST INFRASTRUCTURE; PARITY UNPINNED (no SW in the reference).
 * Full-matrix, 32-bit, O(L*W) memory: a direct transcription of SPEC.md sections 2-6. */
#include "sw_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <pthread.h>

#define NEG_INF (INT_MIN / 4)

/* SPEC section 2. N never mismatches: CompareToReference.java:92. */
static inline int32_t sub_score(uint8_t a, uint8_t b, const swo_params* p) {
  if (a >= SWO_BASE_N || b >= SWO_BASE_N) return 0;
  return a == b ? p->match : p->mismatch;
}

static inline int32_t max2(int32_t a, int32_t b) { return a > b ? a : b; }

static int params_ok(const swo_params* p) {
  return p && p->match >= 1 && p->mismatch <= 0 && p->gap_open >= 0 && p->gap_ext >= 1;
}

int swo_align(const uint8_t* q, int L, const uint8_t* r, int W, const swo_params* p,
              swo_result* res, uint32_t* cigar, uint32_t cigar_cap) {
  if (!params_ok(p) || !res || L < 0 || W < 0) return -2;
  memset(res, 0, sizeof(*res));
  if (L == 0 || W == 0) { res->flags = SWO_FLAG_UNALIGNED; return 0; }

  const size_t stride = (size_t)W + 1;
  const size_t cells = ((size_t)L + 1) * stride;
  int32_t* H = (int32_t*)malloc(cells * sizeof(int32_t));
  int32_t* E = (int32_t*)malloc(cells * sizeof(int32_t));
  int32_t* F = (int32_t*)malloc(cells * sizeof(int32_t));
  if (!H || !E || !F) { free(H); free(E); free(F); return -2; }
  const int32_t go = p->gap_open, ge = p->gap_ext;

  /* SPEC section 3 boundaries */
  for (int j = 0; j <= W; ++j) { H[j] = 0; E[j] = NEG_INF; F[j] = NEG_INF; }
  for (int i = 1; i <= L; ++i) { H[i * stride] = 0; E[i * stride] = NEG_INF; F[i * stride] = NEG_INF; }

  /* SPEC section 3 recurrence + section 4 arg-max (smallest j, then smallest i: positional, so scan and compare keys) */
  int32_t S = 0; int ie = 0, je = 0;
  for (int i = 1; i <= L; ++i) {
    for (int j = 1; j <= W; ++j) {
      const size_t c = (size_t)i * stride + j;
      const int32_t e = max2(E[c - 1], H[c - 1] - go) - ge;
      const int32_t f = max2(F[c - stride], H[c - stride] - go) - ge;
      int32_t h = H[c - stride - 1] + sub_score(q[i - 1], r[j - 1], p);
      h = max2(h, e); h = max2(h, f); h = max2(h, 0);
      E[c] = e; F[c] = f; H[c] = h;
      if (h > S || (h == S && h > 0 && (j < je || (j == je && i < ie)))) { S = h; ie = i; je = j; }
    }
  }
  if (S == 0) { res->flags = SWO_FLAG_UNALIGNED; free(H); free(E); free(F); return 0; }

  /* SPEC section 5 traceback; ops collected back to front, one per base, then run-length encoded */
  uint8_t* ops = (uint8_t*)malloc((size_t)L + (size_t)W + 2);
  if (!ops) { free(H); free(E); free(F); return -2; }
  size_t nops = 0;
  int i = ie, j = je, edits = 0;
  enum { ST_H, ST_F, ST_E } st = ST_H;
  for (;;) {
    const size_t c = (size_t)i * stride + j;
    if (st == ST_H) {
      if (H[c] == 0) break;
      const int32_t sc = sub_score(q[i - 1], r[j - 1], p);
      if (H[c] == H[c - stride - 1] + sc) {
        ops[nops++] = SWO_OP_M;
        if (q[i - 1] < SWO_BASE_N && r[j - 1] < SWO_BASE_N && q[i - 1] != r[j - 1]) ++edits;
        --i; --j;
      } else if (H[c] == F[c]) st = ST_F;
      else st = ST_E;                      /* H[c] == E[c] by construction */
    } else if (st == ST_F) {
      ops[nops++] = SWO_OP_I; ++edits;
      if (i > 1 && F[c] == F[c - stride] - ge) { --i; }
      else { --i; st = ST_H; }
    } else {
      ops[nops++] = SWO_OP_D; ++edits;
      if (j > 1 && E[c] == E[c - 1] - ge) { --j; }
      else { --j; st = ST_H; }
    }
  }
  res->score = S; res->ref_start = j + 1; res->ref_end = je; res->q_start = i + 1; res->q_end = ie;
  res->edit_distance = edits;

  /* SPEC section 6 CIGAR: leading S, merged ops front to back, trailing S */
  uint32_t n = 0; int overflow = 0;
#define PUSH(len, op) do { if ((len) > 0) { if (n < cigar_cap && cigar) cigar[n] = ((uint32_t)(len) << 4) | (op); else overflow = 1; ++n; } } while (0)
  PUSH(res->q_start - 1, SWO_OP_S);
  for (size_t k = nops; k > 0;) {
    const uint8_t op = ops[k - 1]; uint32_t len = 0;
    while (k > 0 && ops[k - 1] == op) { ++len; --k; }
    PUSH(len, op);
  }
  PUSH(L - res->q_end, SWO_OP_S);
#undef PUSH
  res->cigar_n = n;
  free(ops); free(H); free(E); free(F);
  return overflow ? -1 : 0;
}

/* ---- batch driver (pthreads over independent tasks; static interleaved partition) ---- */
typedef struct {
  const uint8_t* q_codes; const uint64_t* q_off; const uint32_t* q_len;
  const uint8_t* r_codes; const uint64_t* r_off; const uint32_t* r_len;
  const uint32_t* pair_ref_idx; uint32_t n; const swo_params* p;
  swo_result* res; uint32_t* cigars; uint32_t cigar_stride; int tid, nthreads, status;
} batch_arg;

static void* batch_worker(void* v) {
  batch_arg* a = (batch_arg*)v;
  for (uint32_t t = (uint32_t)a->tid; t < a->n; t += (uint32_t)a->nthreads) {
    const uint32_t ri = a->pair_ref_idx ? a->pair_ref_idx[t] : t;
    int rc = swo_align(a->q_codes + a->q_off[t], (int)a->q_len[t], a->r_codes + a->r_off[ri], (int)a->r_len[ri],
                       a->p, &a->res[t], a->cigars ? a->cigars + (size_t)t * a->cigar_stride : NULL,
                       a->cigars ? a->cigar_stride : 0);
    if (rc != 0 && a->status == 0) a->status = rc;
  }
  return NULL;
}

int swo_align_batch(const uint8_t* q_codes, const uint64_t* q_off, const uint32_t* q_len,
                    const uint8_t* r_codes, const uint64_t* r_off, const uint32_t* r_len,
                    const uint32_t* pair_ref_idx, uint32_t n, const swo_params* p,
                    swo_result* res, uint32_t* cigars, uint32_t cigar_stride, int nthreads) {
  if (nthreads <= 0) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256]; batch_arg args[256];
  for (int t = 0; t < nthreads; ++t) {
    batch_arg a = { q_codes, q_off, q_len, r_codes, r_off, r_len, pair_ref_idx, n, p, res, cigars, cigar_stride, t, nthreads, 0 };
    args[t] = a;
    if (t > 0) pthread_create(&th[t], NULL, batch_worker, &args[t]);
  }
  batch_worker(&args[0]);
  int status = args[0].status;
  for (int t = 1; t < nthreads; ++t) { pthread_join(th[t], NULL); if (!status) status = args[t].status; }
  return status;
}

/* ---- sequence conventions ---- */

/* Sequence.java:10-13 (A=0 C=1 T=2 G=3); case-folding per CompareToReference.java:90-91; non-ACGT -> N (SPEC section 1). */
uint8_t swo_code_of_char(char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'T': case 't': return 2;
    case 'G': case 'g': return 3;
    default: return SWO_BASE_N;
  }
}
char swo_char_of_code(uint8_t code) { return code < 4 ? "ACTG"[code] : 'N'; }   /* Sequence.java:83-96 */

void swo_encode(const char* s, size_t n, uint8_t* codes) { for (size_t k = 0; k < n; ++k) codes[k] = swo_code_of_char(s[k]); }

/* Sequence.java:31-41: base k -> byte k*2/8, shifted left by (k*2)%8. N goes to the side bit-plane (SPEC section 1). */
void swo_pack2(const uint8_t* codes, size_t n, uint8_t* packed, uint8_t* nmask) {
  memset(packed, 0, (n + 3) / 4);
  if (nmask) memset(nmask, 0, (n + 7) / 8);
  for (size_t k = 0; k < n; ++k) {
    if (codes[k] < SWO_BASE_N) packed[k / 4] |= (uint8_t)(codes[k] << ((k * 2) % 8));
    else if (nmask) nmask[k / 8] |= (uint8_t)(1u << (k % 8));
  }
}
/* Sequence.java:51-72 */
void swo_unpack2(const uint8_t* packed, const uint8_t* nmask, size_t n, uint8_t* codes) {
  for (size_t k = 0; k < n; ++k) {
    uint8_t c = (uint8_t)((packed[k / 4] >> ((k * 2) % 8)) & 3u);
    if (nmask && ((nmask[k / 8] >> (k % 8)) & 1u)) c = SWO_BASE_N;
    codes[k] = c;
  }
}
/* ReverseComplementor.java:15-22 (A<->T, C<->G only), :38-44 (everything else unchanged), :50-55 (reverse then complement). */
void swo_revcomp_text(const char* in, size_t n, char* out) {
  for (size_t k = 0; k < n; ++k) {
    char c = in[n - 1 - k];
    switch (c) { case 'A': c = 'T'; break; case 'T': c = 'A'; break; case 'C': c = 'G'; break; case 'G': c = 'C'; break; default: break; }
    out[k] = c;
  }
}
void swo_revcomp_codes(const uint8_t* in, size_t n, uint8_t* out) {
  static const uint8_t comp[5] = { 2, 3, 0, 1, 4 };        /* A<->T (0<->2), C<->G (1<->3), N stays */
  for (size_t k = 0; k < n; ++k) out[k] = comp[in[n - 1 - k] > 4 ? 4 : in[n - 1 - k]];
}

int32_t swo_rescore(const uint8_t* q, int L, const uint8_t* r, int W, const swo_params* p,
                    const swo_result* res, const uint32_t* cigar, int32_t* edit_out) {
  int qi = 0, rj = res->ref_start - 1, edits = 0; int32_t sc = 0;
  for (uint32_t k = 0; k < res->cigar_n; ++k) {
    const int len = (int)(cigar[k] >> 4); const uint32_t op = cigar[k] & 15u;
    if (op == SWO_OP_S) qi += len;
    else if (op == SWO_OP_M) {
      for (int t = 0; t < len; ++t) {
        if (qi >= L || rj >= W) return INT32_MIN;
        sc += sub_score(q[qi], r[rj], p);
        if (q[qi] < SWO_BASE_N && r[rj] < SWO_BASE_N && q[qi] != r[rj]) ++edits;
        ++qi; ++rj;
      }
    } else if (op == SWO_OP_I) { sc -= p->gap_open + p->gap_ext * len; qi += len; edits += len; }
    else if (op == SWO_OP_D) { sc -= p->gap_open + p->gap_ext * len; rj += len; edits += len; }
    else return INT32_MIN;
  }
  if (qi != L || rj != res->ref_end) return INT32_MIN;
  if (edit_out) *edit_out = edits;
  return sc;
}
