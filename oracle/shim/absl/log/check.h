// Minimal stand-in for absl/log/check.h so that the reference's CPU templates
// (utils/include/*_cpu.hpp) compile without the vendored abseil tree.
// TEST INFRASTRUCTURE ONLY (used by oracle/ref_shim.cu).
#ifndef ORACLE_SHIM_ABSL_LOG_CHECK_H_
#define ORACLE_SHIM_ABSL_LOG_CHECK_H_
#include <cstdlib>
#include <iostream>
namespace oracle_shim {
struct Voidify {
  void operator&(std::ostream&) {}
};
struct Fatal {
  std::ostream& stream() { return std::cerr; }
  ~Fatal() {
    std::cerr << std::endl;
    std::abort();
  }
};
}  // namespace oracle_shim
#define CHECK(cond)                                 \
  (cond) ? (void)0                                  \
         : ::oracle_shim::Voidify() &               \
               ::oracle_shim::Fatal().stream()      \
                   << "Check failed: " #cond " "
#define CHECK_EQ(a, b) CHECK((a) == (b))
#define CHECK_GT(a, b) CHECK((a) > (b))
#endif  // ORACLE_SHIM_ABSL_LOG_CHECK_H_
