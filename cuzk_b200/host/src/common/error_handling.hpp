// common/error_handling.hpp -- exception types and validators of the cuZK host interface.
//
// Stand-alone counterpart of the reference's src/common/error_handling.hpp:15-55 (same names, same base
// classes, same messages) so that code written against the reference keeps compiling when cuzk_b200 is
// built outside the reference tree.  When the host layer is dropped into the reference tree the
// reference's own header is used instead.  The CUDA_CHECK_* macros of the reference
// (:59-111) have no counterpart: no CUDA runtime call crosses the C ABI.
#pragma once

#include <cstddef>
#include <stdexcept>
#include <string>

namespace cuZK {
namespace ErrorHandling {

struct ValidationError : std::invalid_argument {
  explicit ValidationError(const std::string &what) : std::invalid_argument(what) {}
};
struct ComputationError : std::runtime_error {
  explicit ComputationError(const std::string &what) : std::runtime_error(what) {}
};
struct IndexError : std::out_of_range {
  explicit IndexError(const std::string &what) : std::out_of_range(what) {}
};

inline void validate_range(size_t value, size_t lo, size_t hi, const std::string &name) {
  if (value >= lo && value <= hi) return;
  throw ValidationError(name + " must be between " + std::to_string(lo) + " and " + std::to_string(hi) + ", got " +
                        std::to_string(value));
}
inline void validate_index(size_t index, size_t bound, const std::string &context) {
  if (index < bound) return;
  throw IndexError(context + ": index " + std::to_string(index) + " out of range [0, " + std::to_string(bound) + ")");
}
inline void validate_non_empty(size_t size, const std::string &context) {
  if (size == 0) throw ValidationError(context + " cannot be empty");
}

}  // namespace ErrorHandling
}  // namespace cuZK
