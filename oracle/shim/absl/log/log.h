// Minimal stand-in for absl/log/log.h (see check.h in this directory).
#ifndef ORACLE_SHIM_ABSL_LOG_LOG_H_
#define ORACLE_SHIM_ABSL_LOG_LOG_H_
#include <iostream>
#define LOG(severity) std::cerr
#endif  // ORACLE_SHIM_ABSL_LOG_LOG_H_
