// monitor.cpp -- logger, counters, exceptions, GPU probe.
#include "trv/monitor.hpp"

#include <chrono>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <filesystem>

#include "trvb.h"

namespace trv {
namespace sys {

int currTask = 0;
double gbytesMem = 0., gbytesMaxMem = 0.;
double gbytesMemGPU = 0., gbytesMaxMemGPU = 0.;
int count_rgrid = 0, count_cgrid = 0;
float count_grid = 0.f;
int max_count_rgrid = 0, max_count_cgrid = 0;
float max_count_grid = 0.f;
int count_fft = 0, count_ifft = 0;

void update_maxmem(bool gpu) {
  if (gpu) {
    if (gbytesMemGPU > gbytesMaxMemGPU) gbytesMaxMemGPU = gbytesMemGPU;
  } else {
    if (gbytesMem > gbytesMaxMem) gbytesMaxMem = gbytesMem;
  }
}

void update_maxcntgrid() {
  if (count_rgrid > max_count_rgrid) max_count_rgrid = count_rgrid;
  if (count_cgrid > max_count_cgrid) max_count_cgrid = count_cgrid;
  if (count_grid > max_count_grid) max_count_grid = count_grid;
}

extern "C" int trvb_device_count(void);   // csrc/trvb_ctx.cu

int get_gpu_count(bool sys) {
  int num = trvb_device_count();
  if (sys) return num;
  // TRV_GPU_MAXNUM caps the devices used (S/monitor.cpp:270-281).
  const char* cap = std::getenv("TRV_GPU_MAXNUM");
  if (cap != nullptr) {
    int c = std::atoi(cap);
    if (c >= 0 && c < num) num = c;
  }
  return num;
}

bool is_gpu_available() { return get_gpu_count(true) > 0; }

bool is_gpu_enabled() {
  if (get_gpu_count() <= 0) return false;
  const char* mode = std::getenv("TRV_GPU_MODE");   // S/monitor.cpp:303-318
  if (mode != nullptr) {
    std::string m(mode);
    if (m == "false" || m == "no" || m == "off" || m == "0") return false;
  }
  return true;
}

bool has_extension(const std::string& fname, const std::string& fext) {
  return fname.size() >= fext.size()
    && fname.compare(fname.size() - fext.size(), fext.size(), fext) == 0;
}

std::vector<std::string> split_string(const std::string& str, const std::string& delimiter) {
  std::vector<std::string> parts;
  if (str.empty()) return parts;
  if (delimiter.empty()) { parts.push_back(str); return parts; }
  size_t from = 0;
  for (;;) {
    const size_t at = str.find(delimiter, from);
    if (at == std::string::npos) break;
    parts.push_back(str.substr(from, at - from));
    from = at + delimiter.size();
  }
  if (from < str.size()) parts.push_back(str.substr(from));
  return parts;
}

void expand_envar_in_path(std::string& path_str) {
  size_t from = 0;
  for (;;) {
    const size_t open = path_str.find("${", from);
    if (open == std::string::npos) return;
    const size_t close = path_str.find('}', open + 2);
    if (close == std::string::npos) return;
    const std::string name = path_str.substr(open + 2, close - open - 2);
    const char* value = name.empty() ? nullptr : std::getenv(name.c_str());
    if (value != nullptr) {
      path_str.replace(open, close - open + 1, value);
      from = open + std::strlen(value);
    } else {
      from = close + 1;   // unset: left as written
    }
  }
}

bool if_filepath_is_set(const std::string& pathstr) {
  // a path is "set" when it is not blank and does not name a directory (trailing '/')
  if (pathstr.empty() || pathstr.back() == '/') return false;
  for (char c : pathstr) if (!std::isspace(static_cast<unsigned char>(c))) return true;
  return false;
}

void make_write_dir(std::string dirstr) {
  if (dirstr.empty() || dirstr == "." || dirstr == "./" || dirstr == "/") return;
  while (!dirstr.empty() && dirstr.back() == '/') dirstr.pop_back();
  std::error_code ec;
  const bool created = std::filesystem::create_directories(std::filesystem::path(dirstr), ec);
  if (created) {
    if (currTask == 0) logger.info("Directory created: %s", dirstr.c_str());
  } else if (ec) {
    if (currTask == 0) logger.error("Failed to create directory: %s", dirstr.c_str());
    throw IOError("Failed to create directory: %s", dirstr.c_str());
  }
}

bool is_colourable() {
  // colour only on request (TRV_INTERACTIVE) and on a terminal that advertises it
  const char* term = std::getenv("TERM");
  const char* inter = std::getenv("TRV_INTERACTIVE");
  if (term == nullptr || inter == nullptr || std::strstr(term, "color") == nullptr) return false;
  const std::string v(inter);
  return v == "true" || v == "yes" || v == "on" || v == "1";
}

void exit_fatal(const std::string& msg) {
  std::printf(is_colourable() ? "\n\033[1;37;41mFATAL\033[0m: %s\n" : "\nFATAL: %s\n", msg.c_str());
  std::fflush(stdout);
  std::exit(EXIT_FAILURE);
}

void display_help() {
  std::printf(
    "Triumvirate (B200 build): three-point clustering measurements in LSS\n\n"
    "Usage: triumvirate [-h] [-V] <parameter-ini-file>\n\n"
    "Positional arguments:\n"
    "  <parameter-ini-file>  path to the parameter INI file\n\n"
    "Options:\n"
    "  -h, --help     show help message and exit\n"
    "  -V, --version  show version and licensing information and exit\n");
}

void display_prog_logo() {
  std::printf("\n  TRIUMVIRATE  --  Three-Point Clustering Measurements in LSS (B200 device build)\n\n");
}

void display_prog_licence(bool brief) {
  std::printf("Estimator definitions after Triumvirate (Wang & Sugiyama), GPL-3.0-or-later.\n");
  if (!brief) {
    std::printf(
      "This program is free software: you can redistribute it and/or modify it under the\n"
      "terms of the GNU General Public License, version 3 or later.  It comes WITHOUT ANY\n"
      "WARRANTY; see the licence text for details.\n");
  }
  std::printf("\n");
}

void display_prog_info(bool runtime) {
  std::printf(runtime ? "RUNTIME INFORMATION >\n\n" : "PROGRAM INFORMATION >\n\n");
  std::printf("Device layer: %s\n", trvb_version());
  std::printf("CUDA devices visible / usable: %d / %d\n", get_gpu_count(true), get_gpu_count());
  std::printf("GPU mode: %s (no CPU fallback)\n\n", is_gpu_enabled() ? "enabled" : "DISABLED");
}

void display_prog_logbars(int endpoint) {
  if (endpoint == 0) {
    std::printf("PROGRAM LOG >\n\n");
  } else if (endpoint != 1) {
    throw InvalidParameterError("Invalid endpoint for log bars: %d.", endpoint);
  }
}

Logger logger(INFO);

void Logger::log(int level, const char* tag, const char* fmt, va_list args) {
  if (level < level_limit) return;
  char buf[4096];
  std::vsnprintf(buf, sizeof(buf), fmt, args);
  std::time_t now = std::time(nullptr);
  char ts[32];
  std::strftime(ts, sizeof(ts), "%Y-%m-%d %H:%M:%S", std::localtime(&now));
  std::printf("[%s %s C++] %s\n", ts, tag, buf);
}

#define TRV_LOG_METHOD(NAME, LEVEL, TAG)          \
  void Logger::NAME(const char* fmt, ...) {       \
    va_list args;                                 \
    va_start(args, fmt);                          \
    log(LEVEL, TAG, fmt, args);                   \
    va_end(args);                                 \
  }
TRV_LOG_METHOD(debug, DBUG, "DBUG")
TRV_LOG_METHOD(stat, STAT, "STAT")
TRV_LOG_METHOD(info, INFO, "INFO")
TRV_LOG_METHOD(warn, WARN, "WARN")
TRV_LOG_METHOD(error, ERRO, "ERRO")
#undef TRV_LOG_METHOD

#define TRV_DEFINE_ERROR(NAME, BASE)                               \
  NAME::NAME(const char* fmt_string, ...) : BASE("") {             \
    char buf[4096];                                                \
    va_list args;                                                  \
    va_start(args, fmt_string);                                    \
    std::vsnprintf(buf, sizeof(buf), fmt_string, args);            \
    va_end(args);                                                  \
    err_mesg = buf;                                                \
  }                                                                \
  const char* NAME::what() const noexcept { return err_mesg.c_str(); }
TRV_DEFINE_ERROR(UnimplementedError, std::logic_error)
TRV_DEFINE_ERROR(IOError, std::runtime_error)
TRV_DEFINE_ERROR(InvalidParameterError, std::invalid_argument)
TRV_DEFINE_ERROR(InvalidDataError, std::runtime_error)
TRV_DEFINE_ERROR(DeviceError, std::runtime_error)
#undef TRV_DEFINE_ERROR

}  // namespace sys
}  // namespace trv
