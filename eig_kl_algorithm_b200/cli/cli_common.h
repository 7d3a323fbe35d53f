// cli_common.h -- helpers shared by the three drop-in executables
#pragma once
#include <sys/stat.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <string>
#include "eigkl.h"

inline void create_dir(const std::string &name) {       // cKL.cpp:408-417, cEIG.cpp:68-75
  struct stat info;
  if (stat(name.c_str(), &info) != 0) mkdir(name.c_str(), 0755);
}
inline std::string base_name(const std::string &path) { // cEIG.cpp:78-80, cKL.cpp:419-422
  return std::filesystem::path(path).filename().string();
}
// EIGKL_TIMING=1: wall clock since process start after each stage, on stderr (where a drop-in's seconds go)
struct StageClock {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(), last = t0;
  bool on = getenv("EIGKL_TIMING") != nullptr;
  void tick(const char *what) {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[timing] %-28s +%8.3f ms  (total %9.3f ms)\n", what, std::chrono::duration<double, std::milli>(now - last).count(),
            std::chrono::duration<double, std::milli>(now - t0).count());
    last = now;
  }
};
inline int device_from_env() {
  const char *e = getenv("EIGKL_DEVICE");
  return e ? atoi(e) : 0;
}
