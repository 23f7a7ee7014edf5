/*
 * cuzk_b200.h -- C ABI of the B200-native cuZK hot path (libcuzk_b200.so).
 *
 * The reference (davencyw/cuZK) has no C ABI: its boundary is a set of C++ classes that tests
 * and benchmarks include directly.  Each entry point below names the reference interface it
 * replaces (file:line under the reference tree).  The C++ host classes with the reference's own
 * names (cuzk_b200/host/) are thin wrappers over these functions; INTEGRATION.md shows the
 * binding a reference maintainer would add.
 *
 * Conventions
 *   - Field elements: 4 x uint64_t little-endian limbs, plain canonical form, 32 bytes, i.e. the
 *     memory layout of `struct FieldElement` (src/poseidon/field_arithmetic.hpp:11-14) and of
 *     `CudaFieldElement` (src/poseidon/cuda/cuda_field_element.cuh:13-113).  Arrays are packed AoS.
 *   - `mem`: CUZK_MEM_DEVICE = every pointer is device memory on the current CUDA device and the
 *     call is asynchronous on `stream`; CUZK_MEM_HOST = every pointer is host memory, the library
 *     stages through its own device buffers and returns after the results are in the host buffer.
 *   - Devices: a process may initialise several devices (cuzk_init(d) for each); all library state
 *     (constants, padding roots, staging buffers, internal streams) is kept per device.  Every entry
 *     point runs on the CALLING THREAD'S CURRENT CUDA device, which must be an initialised one.  One
 *     convenience: a thread whose current device was never initialised while exactly one other device
 *     was (worker threads of a single-GPU program that never call cudaSetDevice) is switched to that
 *     device for the duration of the call.  Tree handles and cuzk_mg_* handles remember their devices
 *     and run there whatever the caller's current device is.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - Return value: CUZK_OK (0) or a negative error code; cuzk_last_error() gives a message
 *     (thread-local).  Nothing throws, nothing calls exit, there is NO CPU fallback: without a
 *     usable GPU every compute call fails with CUZK_ERR_CUDA.
 *   - Results are bit-exact with the reference CPU implementation (including its non-standard
 *     512->256-bit reduction); see DESIGN.md.
 */
#ifndef CUZK_B200_H
#define CUZK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUZK_OK 0
#define CUZK_ERR_INVALID (-1) /* bad argument (arity outside 2..8, NULL pointer with n > 0, ...) */
#define CUZK_ERR_CUDA (-2)    /* CUDA runtime error, no device, or library not initialised */
#define CUZK_ERR_CONSTANTS (-3) /* generated round constants do not fit the compiled-in fast path */

#define CUZK_MEM_DEVICE 0
#define CUZK_MEM_HOST 1

/* element-wise field operations, reference: CudaFieldArithmetic::batch_{add,subtract,multiply,square,power5}
 * (src/poseidon/cuda/field_arithmetic_cuda.cuh:34-51) */
#define CUZK_FR_ADD 0
#define CUZK_FR_SUB 1
#define CUZK_FR_MUL 2
#define CUZK_FR_SQR 3
#define CUZK_FR_POW5 4

/* ---- lifecycle: CudaFieldArithmetic::initialize/cleanup (field_arithmetic_cuda.cuh:28-31),
 *      CudaPoseidonHash ctor/dtor (poseidon_cuda.cuh:26-27), CudaNaryMerkleTree::initialize_cuda/cleanup_cuda
 *      (merkle_tree_cuda.cuh:101-102).  Reference-counted and idempotent; never resets the device. ---- */
int cuzk_init(int device);      /* also makes `device` the calling thread's current device, like the reference's cudaSetDevice(0) */
int cuzk_shutdown(void);        /* releases the calling thread's current device (or the only initialised one) */
int cuzk_is_initialized(void);  /* on any device */
int cuzk_is_initialized_on(int device);
int cuzk_device_count(void); /* CudaFieldArithmetic::get_device_count (field_arithmetic_cuda.cuh:60) */
/* device properties for CudaFieldArithmetic::print_device_info (field_arithmetic_cuda.cu:633-645) and
 * CudaMerkleUtils::check_cuda_compatibility (merkle_tree_cuda.cu:596-616), so host code needs no CUDA runtime */
typedef struct cuzk_device_info {
  char name[256];
  int cc_major, cc_minor;
  int sm_count;
  int max_threads_per_block;
  size_t total_mem_bytes;
} cuzk_device_info_t;
int cuzk_device_info(int device, cuzk_device_info_t *out);
const char *cuzk_last_error(void);
const char *cuzk_version(void);
/* number of kernels this library has launched since load (bench.py reports it as gpu_launches) */
uint64_t cuzk_launch_count(void);

/* ---- Fr batch ops (out[i] = op(a[i], b[i]); b ignored for SQR/POW5).  Inputs may be any 256-bit values. ---- */
int cuzk_fr_batch(int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n, int mem, void *stream);

/* ---- Poseidon batch hashing: IPoseidonCudaHash (src/poseidon/cuda/poseidon_interface_cuda.hpp:27-47) ---- */
/* batch_hash_single (:32-33): out[i] = PoseidonHash::hash_single(in[i])   (poseidon.cpp:89-91) */
int cuzk_poseidon_hash_single(const uint64_t *in, uint64_t *out, size_t n, int mem, void *stream);
/* batch_hash_pairs (:35-37): out[i] = PoseidonHash::hash_pair(left[i], right[i])   (poseidon.cpp:93-96) */
int cuzk_poseidon_hash_pairs(const uint64_t *left, const uint64_t *right, uint64_t *out, size_t n, int mem,
                             void *stream);
/* batch_permutation (:39): in-place PoseidonHash::permutation on n states of 3 elements (poseidon.cpp:60-87) */
int cuzk_poseidon_permutation(uint64_t *states, size_t n, int mem, void *stream);
/* out[i] = PoseidonHash::sponge(in[i*width .. i*width+width-1], FieldElement(domain_sep))  (poseidon.cpp:103-126);
 * domain_sep 3 = hash_multiple (:98-101) = device_hash_n (src/poseidon/cuda/poseidon_cuda.cu:118-142). width <= 2^20. */
int cuzk_poseidon_sponge(const uint64_t *in, size_t width, uint64_t domain_sep, uint64_t *out, size_t n, int mem,
                         void *stream);
/* round constants (192 x 4 limbs) and MDS matrix (9 x 4 limbs) as the library uses them; host output.
 * PoseidonConstants::ROUND_CONSTANTS / MDS_MATRIX (src/poseidon/poseidon.hpp:19-40) */
int cuzk_poseidon_constants(uint64_t *round_constants_out, uint64_t *mds_out);

/* ---- Merkle geometry (pure host arithmetic) ---- */
/* leaf level padded to arity^L >= n (merkle_tree.cpp:50-53) */
size_t cuzk_merkle_padded_leaves(size_t n, unsigned arity);
/* number of level arrays, leaf level and root included; 0 for n == 0 (integer loop of merkle_tree.cpp:66-97) */
size_t cuzk_merkle_num_levels(size_t n, unsigned arity);
/* sum of the sizes of all padded levels */
size_t cuzk_merkle_total_nodes(size_t n, unsigned arity);
/* NaryMerkleTree::calculate_tree_height (merkle_tree.cpp:359-367): the reference's floating-point
 * formula, for the get_tree_height() getter only -- never used to size anything */
size_t cuzk_merkle_tree_height(size_t leaf_count, unsigned arity);

/* NaryMerkleTree::compute_empty_hash (merkle_tree.cpp:347-357), computed on the GPU; host output */
int cuzk_merkle_empty_hash(unsigned arity, uint64_t out[4]);

/* CudaNaryMerkleTree::build_tree (src/merkle_tree/merkle_tree_cuda.cuh:67, merkle_tree_cuda.cu:141-259) /
 * NaryMerkleTree::build_tree (merkle_tree.cpp:29-100).  Writes every level, level-major:
 * level 0 = the n leaves (un-hashed) padded with empty_hash to padded_leaves, ..., last = root;
 * levels_out holds cuzk_merkle_total_nodes(n, arity) elements.  n >= 1. */
int cuzk_merkle_build(const uint64_t *leaves, size_t n, unsigned arity, uint64_t *levels_out, int mem,
                      void *stream);

/* CudaNaryMerkleTree::build_batch_trees (merkle_tree_cuda.cuh:79-81; a serial loop of full builds in the reference,
 * merkle_tree_cuda.cu:467-482) for `num_trees` trees of `n` leaves each as ONE forest pass: one launch per level (or per
 * two levels) covers every tree.  leaves: num_trees x n elements, tree-major; levels_out: num_trees flat level-major
 * trees of cuzk_merkle_total_nodes(n, arity) elements each, tree-major. */
int cuzk_merkle_build_batch(const uint64_t *leaves, size_t n, size_t num_trees, unsigned arity, uint64_t *levels_out,
                            int mem, void *stream);

/* Subtree-root pass (the multi-GPU shard step, and the "only roots reach HBM" build): the `n` given leaves
 * are the first leaves of `count` consecutive subtrees of arity^height padded leaves each
 * (n <= count * arity^height; the tail is virtual padding).  roots_out[i] = root of subtree i, where
 * all-padding subtrees get the level-`height` padding constant.  No intermediate level is stored. */
int cuzk_merkle_subtree_roots(const uint64_t *leaves, size_t n, unsigned arity, unsigned height, size_t count,
                              uint64_t *roots_out, int mem, void *stream);

/* Root of the tree whose level-`base_level` nodes are given (`count` = a power of arity of them; all of them real):
 * hashes upward with hash_multiple and writes the single root.  Used for the top levels above the per-GPU
 * subtree roots. */
int cuzk_merkle_top_root(const uint64_t *nodes, size_t count, unsigned arity, uint64_t *root_out, int mem,
                         void *stream);

/* the constant root of an all-padding subtree of `height` levels: E_0 = empty_hash(arity),
 * E_{l+1} = hash_multiple(arity copies of E_l); host output */
int cuzk_merkle_padding_root(unsigned arity, unsigned height, uint64_t out[4]);

/* CudaNaryMerkleTree::generate_batch_proofs (merkle_tree_cuda.cuh:74, merkle_tree_cuda.cu:261-339) /
 * NaryMerkleTree::generate_proof (merkle_tree.cpp:113-211) on the level arrays produced by cuzk_merkle_build.
 * For proof q of leaf indices[q] (< n) and level l (0 = leaf level, num_levels-1 of them):
 *   positions_out[q*L + l]                        = (indices[q] / arity^l) mod arity      (MerkleProof::indices)
 *   siblings_out[(q*L + l)*(arity-1) + s]         = the s-th other child of that group     (MerkleProof::path)
 * An index >= n yields positions 0xFFFFFFFF for that proof (the reference returns std::nullopt). */
int cuzk_merkle_prove_batch(const uint64_t *levels, size_t n, unsigned arity, const uint64_t *indices,
                            size_t num_proofs, uint64_t *siblings_out, uint32_t *positions_out, int mem,
                            void *stream);

/* CudaNaryMerkleTree::verify_batch_proofs (merkle_tree_cuda.cuh:75-76, merkle_tree_cuda.cu:341-465) /
 * NaryMerkleTree::verify_proof (merkle_tree.cpp:214-254): results_out[q] = 1 when folding leaf q up
 * `levels` levels with hash_multiple reproduces `root`, 0 otherwise (also 0 when a position >= arity). */
int cuzk_merkle_verify_batch(const uint64_t *leaf_values, const uint64_t *siblings, const uint32_t *positions,
                             size_t levels, unsigned arity, const uint64_t *root, uint8_t *results_out,
                             size_t num_proofs, int mem, void *stream);

/* ---- device-resident tree handle (no reference counterpart: the reference keeps every level in host vectors,
 * merkle_tree_cuda.cuh:43, and re-uploads them per call).  The level arrays stay in HBM; proofs are generated and verified
 * against them without the tree ever crossing PCIe again.  Same layouts as the flat calls above. ---- */
typedef struct cuzk_tree cuzk_tree_t;
/* CudaNaryMerkleTree::build_tree; n >= 1; `mem` describes `leaves` */
int cuzk_tree_build(const uint64_t *leaves, size_t n, unsigned arity, int mem, void *stream, cuzk_tree_t **tree_out);
/* returns the levels to the memory pool in the order of the stream the tree was built on (no device-wide synchronisation);
 * work that uses the tree on OTHER streams must have completed */
int cuzk_tree_free(cuzk_tree_t *tree);
size_t cuzk_tree_leaf_count(const cuzk_tree_t *tree);
size_t cuzk_tree_num_levels(const cuzk_tree_t *tree);
size_t cuzk_tree_total_nodes(const cuzk_tree_t *tree);
unsigned cuzk_tree_arity(const cuzk_tree_t *tree);
int cuzk_tree_device(const cuzk_tree_t *tree);
/* device pointer to the level-major node array (valid until cuzk_tree_free) */
const uint64_t *cuzk_tree_device_levels(const cuzk_tree_t *tree);
/* get_root_hash (merkle_tree_cuda.cuh:84): copies the 32-byte root to host or device memory */
int cuzk_tree_root(const cuzk_tree_t *tree, uint64_t *root_out, int mem, void *stream);
/* get_tree_levels (merkle_tree_cuda.cuh:89): copies every level out, level-major */
int cuzk_tree_levels(const cuzk_tree_t *tree, uint64_t *levels_out, int mem, void *stream);
/* one level (0 = padded leaves ... num_levels-1 = root), padded_leaves / arity^level elements */
int cuzk_tree_level(const cuzk_tree_t *tree, size_t level, uint64_t *level_out, int mem, void *stream);
/* generate_batch_proofs / verify_batch_proofs against the tree's own levels and root; layouts of cuzk_merkle_prove_batch */
int cuzk_tree_prove_batch(const cuzk_tree_t *tree, const uint64_t *indices, size_t num_proofs, uint64_t *siblings_out,
                          uint32_t *positions_out, int mem, void *stream);
int cuzk_tree_verify_batch(const cuzk_tree_t *tree, const uint64_t *leaf_values, const uint64_t *siblings,
                           const uint32_t *positions, uint8_t *results_out, size_t num_proofs, int mem, void *stream);
/* NaryMerkleTree::update_leaf (merkle_tree.cpp:294-301; the reference rebuilds the whole tree per update) for a batch:
 * writes values[q] at leaf indices[q] and re-hashes only the ancestors, one launch per level: count x (levels-1) x
 * ceil(arity/2) permutations instead of a full rebuild.  The tree afterwards equals a serial loop of update_leaf calls in
 * batch order: when an index occurs more than once its LAST value wins.  An index >= leaf_count (the reference throws
 * std::out_of_range, :296-298): host-buffer calls are refused as a whole with CUZK_ERR_INVALID and write nothing;
 * device-pointer calls are asynchronous, so they skip such entries and count them in cuzk_tree_oob_count. */
int cuzk_tree_update_leaves(cuzk_tree_t *tree, const uint64_t *indices, const uint64_t *values, size_t count, int mem,
                            void *stream);
/* how many out-of-range indices device-pointer updates have skipped since the tree was built (a blocking read) */
int cuzk_tree_oob_count(const cuzk_tree_t *tree, uint64_t *count_out);

/* NaryMerkleTree::insert_leaf (merkle_tree.cpp:290-293; a full rebuild per leaf in the reference) for a batch: appends
 * values[0..count) after the last leaf.  While the padded leaf level has room only the new leaves' ancestors are re-hashed;
 * when it overflows, the larger tree is rebuilt on the device from the leaves already there.  Either way the tree equals a
 * fresh build over all leaves. */
int cuzk_tree_append_leaves(cuzk_tree_t *tree, const uint64_t *values, size_t count, int mem, void *stream);

/* ---- several GPUs of one box (SURVEY.md section 8e / 8b "cuzk_mg_*"; no reference counterpart: the reference is single-GPU,
 * field_arithmetic_cuda.cu:20).  Leaves shard into contiguous subtrees, one block of subtrees per GPU; every GPU keeps ALL levels
 * of its subtrees in HBM and serves proofs from them; only the 32-byte subtree roots cross NVLink, in ONE NCCL all-gather issued
 * by the library, and every GPU hashes the few top levels itself.  Batch hashing / verification: contiguous slices, no
 * collective.  NCCL (libnccl.so.2) is loaded at the first cuzk_mg_init_* call, never at library load.
 *   cuzk_mg_init_local : ONE process drives `ngpus` devices (devices == NULL: 0..ngpus-1); every call below then spans them.
 *   cuzk_mg_init_rank  : one process PER GPU (torchrun / MPI): rank 0 calls cuzk_mg_unique_id, the host program carries the
 *                        128 bytes to the other ranks, every rank calls cuzk_mg_init_rank; tree builds are then collective. ---- */
typedef struct cuzk_mg cuzk_mg_t;
typedef struct cuzk_mg_tree cuzk_mg_tree_t;
int cuzk_mg_unique_id(uint8_t id_out[128]);
int cuzk_mg_init_local(int ngpus, const int *devices, cuzk_mg_t **mg_out);
int cuzk_mg_init_rank(int rank, int nranks, int device, const uint8_t id[128], cuzk_mg_t **mg_out);
int cuzk_mg_free(cuzk_mg_t *mg);
int cuzk_mg_world(const cuzk_mg_t *mg, int *nranks_out, int *nlocal_out, int *first_rank_out);
int cuzk_mg_device(const cuzk_mg_t *mg, int local);     /* CUDA device of local shard `local` */
void *cuzk_mg_stream(const cuzk_mg_t *mg, int local);   /* the cudaStream_t the handle's work on that device is ordered on */
int cuzk_mg_nccl_version(void);                         /* of the NCCL that was loaded; 0 = none */
/* which leaves rank `rank` of `nranks` holds in a sharded tree over n leaves: [first, first + count) (count may be 0) */
int cuzk_mg_shard_leaves(size_t n, unsigned arity, int nranks, int rank, size_t *first_out, size_t *count_out);
/* CudaNaryMerkleTree::build_tree (merkle_tree_cuda.cuh:56) over every GPU of the handle.  mem == CUZK_MEM_DEVICE:
 * local_leaves[r] = the leaves of local shard r (cuzk_mg_shard_leaves of rank first_rank + r) in that device's memory;
 * mem == CUZK_MEM_HOST: local_leaves[0] = the WHOLE leaf array in host memory, each shard uploads its own slice.
 * Returns with the root known on the host.  Collective in one-process-per-GPU mode. */
int cuzk_mg_tree_build(cuzk_mg_t *mg, const uint64_t *const *local_leaves, size_t n, unsigned arity, int mem,
                       cuzk_mg_tree_t **tree_out);
int cuzk_mg_tree_free(cuzk_mg_tree_t *tree);
int cuzk_mg_tree_root(const cuzk_mg_tree_t *tree, uint64_t root_out[4]);
size_t cuzk_mg_tree_num_levels(const cuzk_mg_tree_t *tree);      /* of the whole tree, leaves and root included */
size_t cuzk_mg_tree_leaf_count(const cuzk_mg_tree_t *tree);
size_t cuzk_mg_tree_subtree_height(const cuzk_mg_tree_t *tree);  /* hash levels kept per shard subtree */
/* device pointer to local shard `local`'s levels: num_subtrees trees of nodes_per_subtree elements, each level-major */
const uint64_t *cuzk_mg_tree_shard_levels(const cuzk_mg_tree_t *tree, int local, size_t *first_subtree_out,
                                          size_t *num_subtrees_out, size_t *nodes_per_subtree_out);
/* generate_batch_proofs (merkle_tree_cuda.cuh:74): host memory in and out, the wire format of cuzk_merkle_prove_batch over the
 * levels of the WHOLE tree; every query is routed to the local shard that owns the leaf and served from the levels it keeps.
 * Leaves owned by another process's shard (one-process-per-GPU mode) and indices >= n: position 0xFFFFFFFF on every level. */
int cuzk_mg_tree_prove_batch(const cuzk_mg_tree_t *tree, const uint64_t *indices, size_t num_proofs, uint64_t *siblings_out,
                             uint32_t *positions_out);
/* verify_batch_proofs (merkle_tree_cuda.cuh:76) against the sharded tree's root: host memory, slices over the local devices */
int cuzk_mg_tree_verify_batch(const cuzk_mg_tree_t *tree, const uint64_t *leaf_values, const uint64_t *siblings,
                              const uint32_t *positions, uint8_t *results_out, size_t num_proofs);
/* batch_hash_pairs over the local devices: host memory, contiguous slices, one host thread per device */
int cuzk_mg_poseidon_hash_pairs(const cuzk_mg_t *mg, const uint64_t *left, const uint64_t *right, uint64_t *out, size_t n);

/* ---- synthetic inputs: generated on the device so multi-GiB leaf sets need no upload.
 * element i, limb j = splitmix64(seed, 4*(start+i)+j), top limb masked to 60 bits when canonical != 0;
 * u64 leaves: limb 0 = splitmix64(seed, start+i), other limbs 0.  Device pointers only. ---- */
int cuzk_synth_elements(uint64_t *out, size_t n, uint64_t seed, uint64_t start, int canonical, void *stream);
int cuzk_synth_u64_leaves(uint64_t *out, size_t n, uint64_t seed, uint64_t start, void *stream);

/* test hook (device pointers): applies ONE MDS layer (apply_mds_matrix, poseidon.cpp:148-167) in place to n
 * canonical 3-element states; mode 0 = the production fast path with its exact fallback, 1 = exact path only */
int cuzk_debug_mds_layer(uint64_t *states, size_t n, int mode, void *stream);

/* Launches of at most `units` units (hashes, nodes, proofs, states) run on the cooperative kernels -- one permutation spread
 * over a group of lanes (csrc/coop.cuh), a fraction of the one-thread kernels' latency while the chip is not full -- larger
 * ones on the one-thread-per-unit kernels.  0 = never cooperative.  Among the cooperative launches those of at most
 * `wide` units use sixteen lanes per permutation (lowest latency), the others eight (fewer instructions per unit).
 * Both return the previous threshold (defaults: 4736 and 2368 units, profiles/r02_latency_final.json).  Results are identical
 * on every path. */
size_t cuzk_debug_set_coop_max(size_t units);
size_t cuzk_debug_set_coop_wide_max(size_t wide);
/* Full-tree builds (cuzk_merkle_build / cuzk_tree_build of one tree of >= 2^17 leaves): the tree is cut into `groups` groups
 * of subtrees dealt round-robin over `streams` internal streams (<= 16); inside the groups only levels of at most
 * `group_coop_max` nodes take the cooperative kernels.  groups / streams < 1 leave the value unchanged; groups = 1 builds level
 * by level on the caller's stream.  Defaults: 8 groups, 8 streams, 1184 nodes.  Tuning and tests only. */
void cuzk_debug_set_build_plan(int groups, int streams, size_t group_coop_max);
/* Host-buffer calls that move at most `bytes` (inputs + outputs) run their kernel directly on pinned host memory -- the
 * caller's buffers when they are pinned, pinned bounce copies otherwise -- instead of staging through device buffers
 * (default 1 MiB; 0 = always stage).  Returns the previous value.  Tuning and tests only. */
size_t cuzk_debug_set_direct_max(size_t bytes);

/* how many units (hashes, nodes, proof levels, states) were evaluated a second time on the exact path because the fast
 * path met a comparison its top-word test could not decide or a carry that would have to ripple twice (2 to 6 per million
 * permutations on the one-thread kernels, 10 to 14 on the cooperative ones, on random data); a blocking read */
uint64_t cuzk_debug_fallback_count(void);

/* test hook (device pointers): the fast-path field operations on their own -- op 0 = reduce(a), 1 = multiply(a, b),
 * 2 = square(a), 3 = power5(a) -- with the flag each raises when a top-word comparison was undecided.  The property the
 * kernels rely on, and the tests check on crafted boundary values: flags[i] == 0 implies out[i] is the reference result. */
int cuzk_debug_fast_ops(int op, const uint64_t *a, const uint64_t *b, uint64_t *out, uint32_t *flags, size_t n, void *stream);

/* integer-multiply pipe microbenchmark: runs `iters` rounds of dependent-free IMAD.WIDE chains on the whole
 * chip and returns measured 32x32->64 multiply-adds per second (the roofline denominator); variant selects
 * 0 = IMAD.WIDE.U32, 1 = IMAD (lo), 2 = IMAD.HI, 3 = IMAD.WIDE.U32.X carry chains, 4 = IADD3.X carry chains,
 * 5/6/7 = IMAD.WIDE with 1/2/3 carry-chain adds per multiply (counts the multiplies), 8 = SEL, 9 = DFMA,
 * 10 = IMAD.WIDE.U32 with an immediate multiplier, 11 = multiplier in the constant bank, 12 = immediate-form carry chains,
 * 13 / 15 = IMAD.WIDE with 1 / 2 independent DFMA per multiply, 14 = IMAD.WIDE with 1 low IMAD per multiply (multiplies counted) */
int cuzk_imad_peak(int variant, int iters, double *ops_per_second_out);

#ifdef __cplusplus
}
#endif
#endif /* CUZK_B200_H */
