// TEST INFRASTRUCTURE (oracle/_ref build only).  <gflags/gflags.h> stand-in: a flag is a plain global, so the
// reference's common/global_gflags.cpp compiles unmodified and the shim can set FLAGS_xxh3_128bits_seed.
#pragma once
#include <cstdint>
#include <string>
#define DECLARE_int32(n) extern int32_t FLAGS_##n
#define DECLARE_uint32(n) extern uint32_t FLAGS_##n
#define DECLARE_int64(n) extern int64_t FLAGS_##n
#define DECLARE_uint64(n) extern uint64_t FLAGS_##n
#define DECLARE_bool(n) extern bool FLAGS_##n
#define DECLARE_double(n) extern double FLAGS_##n
#define DECLARE_string(n) extern std::string FLAGS_##n
#define DEFINE_int32(n, v, d) int32_t FLAGS_##n = v
#define DEFINE_uint32(n, v, d) uint32_t FLAGS_##n = v
#define DEFINE_int64(n, v, d) int64_t FLAGS_##n = v
#define DEFINE_uint64(n, v, d) uint64_t FLAGS_##n = v
#define DEFINE_bool(n, v, d) bool FLAGS_##n = v
#define DEFINE_double(n, v, d) double FLAGS_##n = v
#define DEFINE_string(n, v, d) std::string FLAGS_##n = v
