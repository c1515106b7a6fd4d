// monitor.hpp -- logger, counters and exception types (thin mirror of the
// reference's trv::sys surface, I/monitor.hpp:249-271,435-565,683-773).
#ifndef TRV_B200_MONITOR_HPP_
#define TRV_B200_MONITOR_HPP_

#include <cstdarg>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

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

// -- paths and strings (I/monitor.hpp:208-245, 830; S/monitor.cpp:38-90, 1012-1040) --
/// Delimiter between the files of a multi-file catalogue.
const std::string fn_delimiter = "::";
/// True when `fname` ends in `fext`.
bool has_extension(const std::string& fname, const std::string& fext);
/// Pieces of `str` between occurrences of `delimiter` (no empty trailing piece).
std::vector<std::string> split_string(const std::string& str, const std::string& delimiter);
/// Replace every ${NAME} by the value of the environment variable NAME (kept when unset).
void expand_envar_in_path(std::string& path_str);
/// I/io.hpp:55-62 (declared in io.hpp by the reference; defined with the I/O code).
bool if_filepath_is_set(const std::string& pathstr);
void make_write_dir(std::string dirstr);

// -- program display and termination (I/monitor.hpp:429, 575, 785-818) --
bool is_colourable();
[[noreturn]] void exit_fatal(const std::string& msg);
void display_help();
void display_prog_logo();
void display_prog_licence(bool brief = false);
void display_prog_info(bool runtime = false);
void display_prog_logbars(int endpoint);

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
