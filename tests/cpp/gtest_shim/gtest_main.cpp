// main() for test binaries built against the gtest shim (TEST INFRASTRUCTURE).
#define GTEST_SHIM_MAIN
#include "gtest/gtest.h"
