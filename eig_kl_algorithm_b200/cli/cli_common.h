// cli_common.h -- helpers shared by the three drop-in executables
#pragma once
#include <sys/stat.h>
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
inline int device_from_env() {
  const char *e = getenv("EIGKL_DEVICE");
  return e ? atoi(e) : 0;
}
