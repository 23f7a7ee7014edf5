/*
 * cuzk_oracle.c -- CPU restatement of the davencyw/cuZK hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This file is the parity oracle for cuzk_b200.  It is NOT part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product path (cuzk_b200/csrc) never links or calls it.
 *
 * Parity is PINNED: oracle/Makefile also compiles the reference's own CPU sources
 * (unmodified, from /root/reference) into oracle/_ref/libcuzk_ref.so, and
 * tests/test_oracle.py (test_oracle_vs_reference_*) + tests/golden/ (generated from that library by
 * tests/golden/generate_golden.py) check every function below against it.
 *
 * Each function cites the reference file:line whose behaviour it restates.
 * Elements are 4 x uint64_t little-endian limbs, plain (non-Montgomery) form,
 * exactly `struct FieldElement` (src/poseidon/field_arithmetic.hpp:11-14).
 *
 * NOTE: the reference "multiply" is deterministic but NOT a true modular multiply
 * (SURVEY.md section 0.1); this file reproduces what the reference computes.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>

typedef unsigned __int128 u128;

/* src/poseidon/field_arithmetic.cpp:12-14 (MODULUS) */
static const uint64_t FR_P[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL,
                                 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
/* src/poseidon/field_arithmetic.cpp:256-258 (k = 2^256 mod p, parsed from hex) */
static const uint64_t FR_K[4] = {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL,
                                 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL};

/* FieldElement::operator< : field_arithmetic.cpp:60-68 (MS limb first) */
static int fr_lt(const uint64_t a[4], const uint64_t b[4]) {
  for (int i = 3; i >= 0; --i) {
    if (a[i] < b[i]) return 1;
    if (a[i] > b[i]) return 0;
  }
  return 0;
}

static int fr_is_zero(const uint64_t a[4]) { return (a[0] | a[1] | a[2] | a[3]) == 0; }

/* subtract_internal : field_arithmetic.cpp:204-219.
 * Quirk kept: b[i] + borrow is computed in 64 bits, so b[i] == 2^64-1 with an
 * incoming borrow wraps to 0 and the borrow is lost. */
static void fr_sub_raw(const uint64_t a[4], const uint64_t b[4], uint64_t r[4]) {
  uint64_t borrow = 0;
  for (int i = 0; i < 4; ++i) {
    uint64_t x = a[i];
    uint64_t y = b[i] + borrow; /* may wrap */
    if (x >= y) {
      r[i] = x - y;
      borrow = 0;
    } else {
      r[i] = x + (UINT64_MAX - y) + 1;
      borrow = 1;
    }
  }
}

/* reduce : field_arithmetic.cpp:244-248 */
void cuzk_oracle_fr_reduce(uint64_t a[4]) {
  while (!fr_lt(a, FR_P)) fr_sub_raw(a, FR_P, a);
}

/* add : field_arithmetic.cpp:172-182 (carry out of limb 3 is dropped, then reduce) */
void cuzk_oracle_fr_add(const uint64_t a[4], const uint64_t b[4], uint64_t r[4]) {
  uint64_t carry = 0, t[4];
  for (int i = 0; i < 4; ++i) {
    u128 s = (u128)a[i] + b[i] + carry;
    t[i] = (uint64_t)s;
    carry = (uint64_t)(s >> 64);
  }
  cuzk_oracle_fr_reduce(t);
  memcpy(r, t, sizeof t);
}

/* subtract : field_arithmetic.cpp:184-202 (no final reduce) */
void cuzk_oracle_fr_sub(const uint64_t a[4], const uint64_t b[4], uint64_t r[4]) {
  uint64_t t[4], out[4];
  if (fr_lt(a, b)) {
    uint64_t carry = 0;
    for (int i = 0; i < 4; ++i) {
      u128 s = (u128)a[i] + FR_P[i] + carry;
      t[i] = (uint64_t)s;
      carry = (uint64_t)(s >> 64);
    }
    fr_sub_raw(t, b, out);
  } else {
    fr_sub_raw(a, b, out);
  }
  memcpy(r, out, sizeof out);
}

/* 4x4 schoolbook, as in multiply()/reduce_512 : field_arithmetic.cpp:223-235 */
static void mul_4x4(const uint64_t a[4], const uint64_t b[4], uint64_t prod[8]) {
  memset(prod, 0, 8 * sizeof(uint64_t));
  for (int i = 0; i < 4; ++i) {
    uint64_t carry = 0;
    for (int j = 0; j < 4; ++j) {
      u128 t = (u128)a[i] * b[j] + prod[i + j] + carry;
      prod[i + j] = (uint64_t)t;
      carry = (uint64_t)(t >> 64);
    }
    prod[i + 4] = carry;
  }
}

/* reduce_512 : field_arithmetic.cpp:250-330 */
void cuzk_oracle_fr_reduce_512(const uint64_t product[8], uint64_t r[4]) {
  const uint64_t *low = product, *high = product + 4;
  uint64_t res[4];
  if (fr_is_zero(high)) { /* :265-269 */
    memcpy(res, low, sizeof res);
    cuzk_oracle_fr_reduce(res);
    memcpy(r, res, sizeof res);
    return;
  }
  uint64_t m1[8];
  mul_4x4(high, FR_K, m1); /* :273-285 */
  uint64_t hc[4];
  memcpy(hc, m1, sizeof hc); /* :298 high_contribution = mult_low */
  if (!fr_is_zero(m1 + 4)) { /* :303 */
    uint64_t m2[8];
    mul_4x4(m1 + 4, FR_K, m2);       /* :305-316 */
    cuzk_oracle_fr_add(hc, m2, hc);  /* :320-322 low half only; add() reduces */
  }
  cuzk_oracle_fr_add(low, hc, res); /* :326 */
  cuzk_oracle_fr_reduce(res);       /* :329 (no-op after add's reduce) */
  memcpy(r, res, sizeof res);
}

/* multiply : field_arithmetic.cpp:221-238 */
void cuzk_oracle_fr_mul(const uint64_t a[4], const uint64_t b[4], uint64_t r[4]) {
  uint64_t prod[8];
  mul_4x4(a, b, prod);
  cuzk_oracle_fr_reduce_512(prod, r);
}

/* square : field_arithmetic.cpp:240-242 */
void cuzk_oracle_fr_sqr(const uint64_t a[4], uint64_t r[4]) { cuzk_oracle_fr_mul(a, a, r); }

/* power5 : field_arithmetic.cpp:332-338 */
void cuzk_oracle_fr_pow5(const uint64_t a[4], uint64_t r[4]) {
  uint64_t a2[4], a4[4], in[4];
  memcpy(in, a, sizeof in);
  cuzk_oracle_fr_sqr(in, a2);
  cuzk_oracle_fr_sqr(a2, a4);
  cuzk_oracle_fr_mul(a4, in, r);
}

/* ---- Poseidon constants : src/poseidon/poseidon.cpp:33-58 ---- */
#define T 3
#define ROUNDS_FULL 8
#define ROUNDS_PARTIAL 56
#define ROUNDS (ROUNDS_FULL + ROUNDS_PARTIAL)

static uint64_t g_rc[ROUNDS * T][4];
static uint64_t g_mds[T * T][4];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void constants_init(void) {
  /* generate_round_constants : poseidon.cpp:33-44 */
  for (size_t i = 0; i < ROUNDS * T; ++i) {
    uint64_t base[4] = {i + 1, 0, 0, 0};
    uint64_t mix[4] = {0x123456789ABCDEFULL, 0, 0, 0};
    uint64_t off[4] = {i * 0x987654321ULL, 0, 0, 0};
    cuzk_oracle_fr_mul(base, mix, g_rc[i]);
    cuzk_oracle_fr_add(g_rc[i], off, g_rc[i]);
  }
  /* generate_mds_matrix : poseidon.cpp:46-58 */
  static const uint64_t m[9] = {7, 23, 8, 26, 5, 4, 15, 20, 9};
  for (int i = 0; i < 9; ++i) {
    g_mds[i][0] = m[i];
    g_mds[i][1] = g_mds[i][2] = g_mds[i][3] = 0;
  }
}

void cuzk_oracle_round_constants(uint64_t out[ROUNDS * T * 4]) {
  pthread_once(&g_once, constants_init);
  memcpy(out, g_rc, sizeof g_rc);
}

void cuzk_oracle_mds(uint64_t out[T * T * 4]) {
  pthread_once(&g_once, constants_init);
  memcpy(out, g_mds, sizeof g_mds);
}

/* permutation : poseidon.cpp:60-87 with add_round_constants :128-134,
 * apply_sbox :136-140, apply_partial_sbox :142-146, apply_mds_matrix :148-167 */
void cuzk_oracle_permutation(uint64_t *state) {
  pthread_once(&g_once, constants_init);
  uint64_t(*s)[4] = (uint64_t(*)[4])state;
  for (int round = 0; round < ROUNDS; ++round) {
    for (int i = 0; i < T; ++i) cuzk_oracle_fr_add(s[i], g_rc[round * T + i], s[i]);
    int full = round < ROUNDS_FULL / 2 || round >= ROUNDS_FULL / 2 + ROUNDS_PARTIAL;
    if (full) {
      for (int i = 0; i < T; ++i) cuzk_oracle_fr_pow5(s[i], s[i]);
    } else {
      cuzk_oracle_fr_pow5(s[0], s[0]);
    }
    uint64_t n[T][4];
    for (int i = 0; i < T; ++i) {
      memset(n[i], 0, sizeof n[i]);
      for (int j = 0; j < T; ++j) {
        uint64_t tmp[4];
        cuzk_oracle_fr_mul(g_mds[i * T + j], s[j], tmp);
        cuzk_oracle_fr_add(n[i], tmp, n[i]);
      }
    }
    memcpy(state, n, sizeof n);
  }
}

/* one apply_mds_matrix (poseidon.cpp:148-167) on n states, in place -- used to test the GPU MDS layer alone */
void cuzk_oracle_batch_mds_layer(uint64_t *states, size_t n) {
  pthread_once(&g_once, constants_init);
  for (size_t k = 0; k < n; ++k) {
    uint64_t(*s)[4] = (uint64_t(*)[4])(states + 12 * k);
    uint64_t nn[T][4];
    for (int i = 0; i < T; ++i) {
      memset(nn[i], 0, sizeof nn[i]);
      for (int j = 0; j < T; ++j) {
        uint64_t tmp[4];
        cuzk_oracle_fr_mul(g_mds[i * T + j], s[j], tmp);
        cuzk_oracle_fr_add(nn[i], tmp, nn[i]);
      }
    }
    memcpy(s, nn, sizeof nn);
  }
}

/* sponge : poseidon.cpp:103-126.  domain separator is a full element. */
void cuzk_oracle_sponge(const uint64_t *inputs, size_t n, const uint64_t ds[4], uint64_t out[4]) {
  uint64_t st[T][4];
  memset(st, 0, sizeof st);
  memcpy(st[0], ds, sizeof st[0]);
  size_t idx = 0;
  while (idx < n) {
    for (int i = 0; i < 2 && idx < n; ++i) {
      cuzk_oracle_fr_add(st[i + 1], inputs + 4 * idx, st[i + 1]);
      idx++;
    }
    cuzk_oracle_permutation(&st[0][0]);
  }
  memcpy(out, st[1], sizeof st[1]);
}

/* hash_single / hash_pair / hash_multiple : poseidon.cpp:89-101 */
void cuzk_oracle_hash_single(const uint64_t in[4], uint64_t out[4]) {
  const uint64_t ds[4] = {1, 0, 0, 0};
  cuzk_oracle_sponge(in, 1, ds, out);
}
void cuzk_oracle_hash_pair(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
  const uint64_t ds[4] = {2, 0, 0, 0};
  uint64_t in[8];
  memcpy(in, l, 32);
  memcpy(in + 4, r, 32);
  cuzk_oracle_sponge(in, 2, ds, out);
}
void cuzk_oracle_hash_multiple(const uint64_t *in, size_t n, uint64_t out[4]) {
  const uint64_t ds[4] = {3, 0, 0, 0};
  cuzk_oracle_sponge(in, n, ds, out);
}

/* ---- batch forms (element-wise loops; layouts = the C-ABI's, include/cuzk_b200.h) ---- */
void cuzk_oracle_batch_fr(int op, const uint64_t *a, const uint64_t *b, uint64_t *r, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    switch (op) {
    case 0: cuzk_oracle_fr_add(a + 4 * i, b + 4 * i, r + 4 * i); break;
    case 1: cuzk_oracle_fr_sub(a + 4 * i, b + 4 * i, r + 4 * i); break;
    case 2: cuzk_oracle_fr_mul(a + 4 * i, b + 4 * i, r + 4 * i); break;
    case 3: cuzk_oracle_fr_sqr(a + 4 * i, r + 4 * i); break;
    case 4: cuzk_oracle_fr_pow5(a + 4 * i, r + 4 * i); break;
    }
  }
}
void cuzk_oracle_batch_hash_single(const uint64_t *in, uint64_t *out, size_t n) {
  for (size_t i = 0; i < n; ++i) cuzk_oracle_hash_single(in + 4 * i, out + 4 * i);
}
void cuzk_oracle_batch_hash_pairs(const uint64_t *l, const uint64_t *r, uint64_t *out, size_t n) {
  for (size_t i = 0; i < n; ++i) cuzk_oracle_hash_pair(l + 4 * i, r + 4 * i, out + 4 * i);
}
void cuzk_oracle_batch_permutation(uint64_t *states, size_t n) {
  for (size_t i = 0; i < n; ++i) cuzk_oracle_permutation(states + 12 * i);
}
/* n hashes of `width` consecutive inputs each, domain separator ds (small integer) */
void cuzk_oracle_batch_sponge(const uint64_t *in, size_t width, uint64_t ds, uint64_t *out, size_t n) {
  const uint64_t d[4] = {ds, 0, 0, 0};
  for (size_t i = 0; i < n; ++i) cuzk_oracle_sponge(in + 4 * width * i, width, d, out + 4 * i);
}

/* multi-threaded pair hashing, used as the "port" CPU baseline when oracle/_ref is absent */
typedef struct { const uint64_t *l, *r; uint64_t *out; size_t lo, hi; } span_t;
static void *pair_worker(void *p) {
  span_t *s = (span_t *)p;
  for (size_t i = s->lo; i < s->hi; ++i) cuzk_oracle_hash_pair(s->l + 4 * i, s->r + 4 * i, s->out + 4 * i);
  return NULL;
}
void cuzk_oracle_batch_hash_pairs_mt(const uint64_t *l, const uint64_t *r, uint64_t *out, size_t n, int threads) {
  pthread_once(&g_once, constants_init);
  if (threads < 1) threads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * threads);
  span_t *sp = (span_t *)malloc(sizeof(span_t) * threads);
  for (int t = 0; t < threads; ++t) {
    sp[t] = (span_t){l, r, out, n * t / threads, n * (t + 1) / threads};
    pthread_create(&th[t], NULL, pair_worker, &sp[t]);
  }
  for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
  free(th);
  free(sp);
}

/* ---- Merkle tree ---- */

/* compute_empty_hash : merkle_tree.cpp:347-357 (hash_multiple of `arity` zeros) */
void cuzk_oracle_empty_hash(size_t arity, uint64_t out[4]) {
  uint64_t z[8 * 4];
  memset(z, 0, sizeof z);
  cuzk_oracle_hash_multiple(z, arity, out);
}

/* calculate_tree_height : merkle_tree.cpp:359-367 (floating-point) */
size_t cuzk_oracle_tree_height_float(size_t leaf_count, size_t arity) {
  if (leaf_count <= 1) return 1;
  return (size_t)ceil(log((double)leaf_count) / log((double)arity)) + 1;
}

/* padded leaf count: merkle_tree.cpp:50-53 */
size_t cuzk_oracle_padded_size(size_t n, size_t arity) {
  size_t p = 1;
  while (p < n) p *= arity;
  return p;
}

/* number of level arrays (leaf level included) for n leaves: loop of merkle_tree.cpp:66-97 */
size_t cuzk_oracle_num_levels(size_t n, size_t arity) {
  if (n == 0) return 0;
  size_t p = cuzk_oracle_padded_size(n, arity), lv = 1;
  while (p > 1) { p /= arity; lv++; }
  return lv;
}

/* total nodes over all levels of the padded tree */
size_t cuzk_oracle_total_nodes(size_t n, size_t arity) {
  if (n == 0) return 0;
  size_t p = cuzk_oracle_padded_size(n, arity), tot = p;
  while (p > 1) { p /= arity; tot += p; }
  return tot;
}

/* build_tree_bottom_up : merkle_tree.cpp:44-100, as flat level arrays.
 * levels_out holds total_nodes elements: level 0 (padded leaves, un-hashed, padding =
 * empty_hash(arity)) first, root last.  n must be >= 1. */
void cuzk_oracle_merkle_build(const uint64_t *leaves, size_t n, size_t arity, uint64_t *levels_out) {
  size_t p = cuzk_oracle_padded_size(n, arity);
  uint64_t e[4];
  cuzk_oracle_empty_hash(arity, e);
  memcpy(levels_out, leaves, n * 32);
  for (size_t i = n; i < p; ++i) memcpy(levels_out + 4 * i, e, 32);
  uint64_t *cur = levels_out;
  while (p > 1) {
    uint64_t *nxt = cur + 4 * p;
    size_t q = p / arity;
    for (size_t i = 0; i < q; ++i) cuzk_oracle_hash_multiple(cur + 4 * arity * i, arity, nxt + 4 * i);
    cur = nxt;
    p = q;
  }
}

/* get_root_hash : merkle_tree.cpp:304-309 (n == 0 -> empty hash) */
void cuzk_oracle_merkle_root(const uint64_t *leaves, size_t n, size_t arity, uint64_t root[4]) {
  if (n == 0) { cuzk_oracle_empty_hash(arity, root); return; }
  size_t tot = cuzk_oracle_total_nodes(n, arity);
  uint64_t *lv = (uint64_t *)malloc(tot * 32);
  cuzk_oracle_merkle_build(leaves, n, arity, lv);
  memcpy(root, lv + 4 * (tot - 1), 32);
  free(lv);
}

/* generate_proof / generate_bottom_up_proof : merkle_tree.cpp:113-211, restated on the
 * level arrays: positions[l] = (index / arity^l) mod arity; siblings[l] = the other
 * arity-1 nodes of that group in index order.  Returns number of proof levels, or -1
 * when leaf_index >= n (reference: std::nullopt). */
long cuzk_oracle_merkle_prove(const uint64_t *levels, size_t n, size_t arity, size_t leaf_index,
                              uint64_t *siblings_out, uint64_t *positions_out) {
  if (n == 0 || leaf_index >= n) return -1;
  size_t p = cuzk_oracle_padded_size(n, arity);
  const uint64_t *cur = levels;
  size_t idx = leaf_index;
  long lv = 0;
  while (p > 1) {
    size_t pos = idx % arity, base = idx - pos, w = 0;
    positions_out[lv] = pos;
    for (size_t c = 0; c < arity; ++c) {
      if (c == pos) continue;
      memcpy(siblings_out + 4 * ((size_t)lv * (arity - 1) + w), cur + 4 * (base + c), 32);
      w++;
    }
    cur += 4 * p;
    p /= arity;
    idx /= arity;
    lv++;
  }
  return lv;
}

/* verify_proof : merkle_tree.cpp:214-254 for a structurally valid proof of `levels`
 * levels; position >= arity -> 0 (reference :228-230). */
int cuzk_oracle_merkle_verify(const uint64_t leaf[4], const uint64_t *siblings, const uint64_t *positions,
                              size_t levels, size_t arity, const uint64_t root[4]) {
  uint64_t cur[4], kids[8 * 4];
  memcpy(cur, leaf, 32);
  for (size_t l = 0; l < levels; ++l) {
    size_t pos = positions[l], w = 0;
    if (pos >= arity) return 0;
    for (size_t c = 0; c < arity; ++c) {
      if (c == pos) memcpy(kids + 4 * c, cur, 32);
      else { memcpy(kids + 4 * c, siblings + 4 * (l * (arity - 1) + w), 32); w++; }
    }
    cuzk_oracle_hash_multiple(kids, arity, cur);
  }
  return memcmp(cur, root, 32) == 0;
}
