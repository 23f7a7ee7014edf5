// field_element.cpp -- value-type members of Poseidon::FieldElement for stand-alone builds of the host layer
// (construction, comparison, text I/O; reference: src/poseidon/field_arithmetic.cpp:25-167).  No arithmetic:
// the CPU arithmetic members are the reference's (the parity oracle) and are not part of this product.
#include "field_arithmetic.hpp"

#include <cstring>
#include <stdexcept>

namespace Poseidon {

FieldElement::FieldElement() : limbs{0, 0, 0, 0} {}
FieldElement::FieldElement(uint64_t value) : limbs{value, 0, 0, 0} {}
FieldElement::FieldElement(uint64_t v0, uint64_t v1, uint64_t v2, uint64_t v3) : limbs{v0, v1, v2, v3} {}
FieldElement::FieldElement(const FieldElement &other) { std::memcpy(limbs, other.limbs, sizeof limbs); }
FieldElement &FieldElement::operator=(const FieldElement &other) {
  if (this != &other) std::memcpy(limbs, other.limbs, sizeof limbs);
  return *this;
}

bool FieldElement::operator==(const FieldElement &other) const { return std::memcmp(limbs, other.limbs, sizeof limbs) == 0; }
bool FieldElement::operator!=(const FieldElement &other) const { return !(*this == other); }
bool FieldElement::operator<(const FieldElement &other) const {
  for (int i = 3; i >= 0; --i)
    if (limbs[i] != other.limbs[i]) return limbs[i] < other.limbs[i];
  return false;
}

std::string FieldElement::to_hex() const {
  static const char digits[] = "0123456789abcdef";
  std::string s(66, '0');
  s[1] = 'x';
  for (int limb = 0; limb < 4; ++limb)
    for (int nib = 0; nib < 16; ++nib) s[65 - (16 * limb + nib)] = digits[(limbs[limb] >> (4 * nib)) & 15];
  return s;
}

std::string FieldElement::to_dec() const {
  if (is_zero()) return "0";
  uint64_t w[4] = {limbs[0], limbs[1], limbs[2], limbs[3]};
  std::string rev;
  while (w[0] | w[1] | w[2] | w[3]) {
    unsigned __int128 rem = 0;  // long division by 10^18 keeps the loop short
    const uint64_t base = 1000000000000000000ull;
    for (int i = 3; i >= 0; --i) {
      unsigned __int128 cur = (rem << 64) | w[i];
      w[i] = (uint64_t)(cur / base);
      rem = cur % base;
    }
    uint64_t chunk = (uint64_t)rem;
    const bool more = (w[0] | w[1] | w[2] | w[3]) != 0;
    for (int d = 0; d < 18 && (more || chunk); ++d) {
      rev.push_back(char('0' + chunk % 10));
      chunk /= 10;
    }
  }
  return std::string(rev.rbegin(), rev.rend());
}

FieldElement FieldElement::from_hex(const std::string &hex) {
  size_t pos = (hex.size() >= 2 && hex[0] == '0' && (hex[1] == 'x' || hex[1] == 'X')) ? 2 : 0;
  const size_t ndigits = hex.size() - pos;
  if (ndigits > 64) throw std::invalid_argument("FieldElement::from_hex: more than 64 hex digits");
  FieldElement r;
  for (size_t k = 0; k < ndigits; ++k) {  // k-th digit from the right
    const char c = hex[hex.size() - 1 - k];
    uint64_t v;
    if (c >= '0' && c <= '9') v = c - '0';
    else if (c >= 'a' && c <= 'f') v = c - 'a' + 10;
    else if (c >= 'A' && c <= 'F') v = c - 'A' + 10;
    else throw std::invalid_argument("FieldElement::from_hex: bad digit");
    r.limbs[k / 16] |= v << (4 * (k % 16));
  }
  return r;
}

bool FieldElement::is_zero() const { return (limbs[0] | limbs[1] | limbs[2] | limbs[3]) == 0; }
void FieldElement::set_zero() { limbs[0] = limbs[1] = limbs[2] = limbs[3] = 0; }

}  // namespace Poseidon
