// monitor.hpp -- logger, counters and exception types (thin mirror of the
// reference's trv::sys surface, I/monitor.hpp:249-271,435-565,683-773).
#ifndef TRV_B200_MONITOR_HPP_
#define TRV_B200_MONITOR_HPP_

#include <cstdarg>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace trv {
namespace sys {

// Program tracking (I/monitor.hpp:249-266).  The mesh lives in HBM, so the
// host "gbytesMem" counters only track the catalogue; gbytesMemGPU tracks
// device meshes.
extern int currTask;
extern double gbytesMem, gbytesMaxMem;
extern double gbytesMemGPU, gbytesMaxMemGPU;
extern int count_rgrid, count_cgrid;
extern float count_grid;
extern int max_count_rgrid, max_count_cgrid;
extern float max_count_grid;
extern int count_fft, count_ifft;

void update_maxmem(bool gpu = false);
void update_maxcntgrid();

template <typename T>
double size_in_gb(long long num) {
  return double(num) * sizeof(T) / (1024. * 1024. * 1024.);
}

// GPU probe (S/monitor.cpp:258-324).  This build has no CPU fallback:
// is_gpu_enabled() being false makes every estimator entry point throw.
int get_gpu_count(bool sys = false);
bool is_gpu_available();
bool is_gpu_enabled();

enum LogLevel { NSET = 0, DBUG = 10, STAT = 20, INFO = 30, WARN = 40, ERRO = 50 };

// Logger with the reference's levels (I/monitor.hpp:435-565).
class Logger {
 public:
  int level_limit;
  explicit Logger(int level = INFO) : level_limit(level) {}
  void reset_level(int level) { level_limit = level; }
  void log(int level, const char* tag, const char* fmt, va_list args);
  void debug(const char* fmt, ...);
  void stat(const char* fmt, ...);
  void info(const char* fmt, ...);
  void warn(const char* fmt, ...);
  void error(const char* fmt, ...);
};

extern Logger logger;

// Exceptions (I/monitor.hpp:683-773).
#define TRV_DECLARE_ERROR(NAME, BASE)                      \
  class NAME : public BASE {                               \
   public:                                                 \
    std::string err_mesg;                                  \
    NAME(const char* fmt_string, ...);                     \
    virtual const char* what() const noexcept;             \
  };
TRV_DECLARE_ERROR(UnimplementedError, std::logic_error)
TRV_DECLARE_ERROR(IOError, std::runtime_error)
TRV_DECLARE_ERROR(InvalidParameterError, std::invalid_argument)
TRV_DECLARE_ERROR(InvalidDataError, std::runtime_error)
// Raised when a device-layer (trvb_*) call fails: CUDA/cuFFT errors are
// never ignored (the reference prints and continues, I/monitor.hpp:162-196).
TRV_DECLARE_ERROR(DeviceError, std::runtime_error)
#undef TRV_DECLARE_ERROR

}  // namespace sys
}  // namespace trv

#endif  // TRV_B200_MONITOR_HPP_
