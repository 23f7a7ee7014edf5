// poseidon/poseidon.hpp -- Poseidon parameters and the CPU hash interface names.
//
// Stand-alone counterpart of the reference's src/poseidon/poseidon.hpp:8-76.  PoseidonParams is what the GPU
// interface needs (STATE_SIZE sizes batch_permutation's state arrays).  The reference's CPU classes PoseidonConstants
// and PoseidonHash (the parity oracle) are absent: cuzk_b200 ships no CPU hash.  benchmark_poseidon* are declared
// because the CPU-vs-GPU benchmark helper calls them; they are defined only by the reference's poseidon.cpp.
#pragma once

#include "field_arithmetic.hpp"

namespace Poseidon {

struct PoseidonParams {
  static constexpr size_t STATE_SIZE = 3;      // t
  static constexpr size_t CAPACITY = 1;        // c
  static constexpr size_t RATE = STATE_SIZE - CAPACITY;
  static constexpr size_t ROUNDS_FULL = 8;     // R_F
  static constexpr size_t ROUNDS_PARTIAL = 56; // R_P
  static constexpr size_t TOTAL_ROUNDS = ROUNDS_FULL + ROUNDS_PARTIAL;
  static constexpr size_t ALPHA = 5;           // S-box exponent
};

struct HashingStats {
  double total_time_ms;
  double avg_time_per_hash_ns;
  size_t hashes_per_second;
  size_t total_hashes;
};

HashingStats benchmark_poseidon(size_t num_iterations = 10000);
HashingStats benchmark_poseidon_pairs(size_t num_pairs = 10000);

}  // namespace Poseidon
