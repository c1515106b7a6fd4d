// monitor.cpp -- logger, counters, exceptions, GPU probe.
#include "trv/monitor.hpp"

#include <chrono>
#include <cstdlib>
#include <ctime>

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
