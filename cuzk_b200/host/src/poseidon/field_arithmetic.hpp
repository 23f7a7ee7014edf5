// poseidon/field_arithmetic.hpp -- the FieldElement value type of the cuZK host interface.
//
// Stand-alone counterpart of the reference's src/poseidon/field_arithmetic.hpp:11-82: the same 4 x 64-bit
// little-endian limb layout (32 bytes, plain canonical form), the same member and namespace names.  Only
// the VALUE-TYPE part is implemented by cuzk_b200 (field_element.cpp: construction, comparison, hex and
// decimal text).  The CPU arithmetic the reference declares next to it (operator+, FieldElement::random,
// FieldArithmetic::*, FieldConstants) is the reference's CPU implementation -- i.e. the parity oracle -- and is
// deliberately absent here: cuzk_b200 has no CPU arithmetic path.  Inside the reference tree the reference's own
// header (a superset of this one, same layout) is the one that gets included; see INTEGRATION.md.
#pragma once

#include <array>
#include <cstdint>
#include <string>
#include <vector>

namespace Poseidon {

struct FieldElement {
  uint64_t limbs[4];  // little-endian: value = sum limbs[i] * 2^(64 i)

  FieldElement();
  explicit FieldElement(uint64_t value);
  FieldElement(uint64_t v0, uint64_t v1, uint64_t v2, uint64_t v3);
  FieldElement(const FieldElement &other);
  FieldElement &operator=(const FieldElement &other);

  bool operator==(const FieldElement &other) const;
  bool operator!=(const FieldElement &other) const;
  bool operator<(const FieldElement &other) const;  // most-significant limb first

  std::string to_hex() const;  // "0x" + 64 lower-case hex digits, most-significant limb first
  std::string to_dec() const;
  static FieldElement from_hex(const std::string &hex);
  bool is_zero() const;
  void set_zero();
};

}  // namespace Poseidon
