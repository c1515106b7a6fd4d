// field.cpp -- device-backed trv::MeshField / trv::FieldStats.
//
// Each method is the reference method of the same name (S/field.cpp) expressed
// as calls into the C-ABI device layer (include/trvb.h).  The mesh stays in
// HBM between calls; the host mirror `field` is refreshed on demand only.
#include "trv/field.hpp"
#include "trv/io.hpp"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>

namespace trvs = trv::sys;
namespace trvm = trv::maths;

namespace trv {

// =====================================================================
// Device helpers
// =====================================================================

namespace dev {

void check(int status, const char* what) {
  if (status != 0) {
    if (trvs::currTask == 0) {
      trvs::logger.error("%s failed: %s", what, trvb_last_error());
    }
    throw trvs::DeviceError("%s failed (status %d): %s", what, status, trvb_last_error());
  }
}

namespace {

typedef std::tuple<int, int, int, int, double, double, double, int> CtxKey;   // device first
std::mutex g_ctx_mutex;
std::map<CtxKey, std::weak_ptr<trvb_ctx> > g_ctx_cache;
std::map<int, std::vector<std::shared_ptr<trvb_ctx> > > g_ctx_recent;   // per device
trvb_ctx* g_last_ctx = nullptr;

}  // namespace

/// The GPU of this process's estimator calls: TRV_GPU_DEVICE, else LOCAL_RANK (one
/// process per GPU under torchrun), else the calling thread's CURRENT CUDA device --
/// so a single-process caller that selected a device (cudaSetDevice,
/// torch.cuda.set_device) and hands over device pointers gets its context there.
/// Ids outside [0, device count) are an error, not wrapped.
int select_device() {
  const int n = trvb_device_count();
  const char* name = "TRV_GPU_DEVICE";
  const char* dev = std::getenv(name);
  if (dev == nullptr || dev[0] == '\0') { name = "LOCAL_RANK"; dev = std::getenv(name); }
  if (dev != nullptr && dev[0] != '\0') {
    char* end = nullptr;
    const long id = std::strtol(dev, &end, 10);
    if (end == dev || *end != '\0' || id < 0 || id >= n) {
      throw trvs::DeviceError("%s=%s does not name one of the %d visible CUDA devices.",
                              name, dev, n);
    }
    return static_cast<int>(id);
  }
  const int cur = trvb_current_device();
  return (cur >= 0 && cur < n) ? cur : 0;
}

DeviceScope::DeviceScope() : prev_(trvb_current_device()) {}
DeviceScope::~DeviceScope() { if (prev_ >= 0) trvb_set_current_device(prev_); }

std::shared_ptr<trvb_ctx> acquire_context(const trv::ParameterSet& params) {
  if (!trvs::is_gpu_enabled()) {
    if (trvs::currTask == 0) {
      trvs::logger.error(
        "No usable CUDA device: triumvirate_b200 has no CPU fallback "
        "(check the driver, TRV_GPU_MODE and TRV_GPU_MAXNUM).");
    }
    throw trvs::DeviceError(
      "No usable CUDA device: triumvirate_b200 has no CPU fallback.");
  }
  // An unset box or mesh passes validate() as in the reference (the program derives them
  // afterwards, S/parameters.cpp:1416-1470); an estimator must not be reached with one.
  for (int ax = 0; ax < 3; ax++) {
    if (!(params.ngrid[ax] > 0) || !(params.boxsize[ax] > 0.)) {
      throw trvs::InvalidParameterError(
        "Mesh is not set: `ngrid` = (%d, %d, %d), `boxsize` = (%lg, %lg, %lg); derive them "
        "with trv::set_boxsize_from_expand / trv::set_ngrid_from_cutoff first.",
        params.ngrid[0], params.ngrid[1], params.ngrid[2],
        params.boxsize[0], params.boxsize[1], params.boxsize[2]);
    }
  }
  const int device = select_device();
  CtxKey key(device, params.ngrid[0], params.ngrid[1], params.ngrid[2],
             params.boxsize[0], params.boxsize[1], params.boxsize[2],
             params.assignment_order);
  std::lock_guard<std::mutex> lock(g_ctx_mutex);
  auto it = g_ctx_cache.find(key);
  if (it != g_ctx_cache.end()) {
    if (auto sp = it->second.lock()) { g_last_ctx = sp.get(); return sp; }
  }
  trvb_ctx* raw = nullptr;
  check(trvb_ctx_create(&raw, device, params.ngrid, params.boxsize,
                        params.assignment_order), "trvb_ctx_create");
  std::shared_ptr<trvb_ctx> sp(raw, [](trvb_ctx* c) { trvb_ctx_destroy(c); });
  g_ctx_cache[key] = sp;
  // Keep the most recent contexts alive between estimator calls so that cuFFT
  // plans, correction tables and the sort scratch are built once per grid
  // (the counterpart of the reference's FFTW wisdom, S/field.cpp:90-169).  Four per
  // device: in single-process multi-GPU mode every device holds its own.
  std::vector<std::shared_ptr<trvb_ctx> >& recent = g_ctx_recent[device];
  recent.push_back(sp);
  if (recent.size() > 4) recent.erase(recent.begin());
  g_last_ctx = raw;
  return sp;
}

trvb_ctx* last_context() { return g_last_ctx; }

namespace {
trvb_comm* g_comm = nullptr;
bool g_prof_enabled() { return profile_enabled(); }
}  // namespace

trvb_comm* process_comm() { return g_comm; }

void comm_unique_id(char id[128]) {
  check(trvb_comm_unique_id(id), "trvb_comm_unique_id");
}

void comm_init(int nranks, int rank, const char id[128]) {
  comm_finalize();
  check(trvb_comm_create(&g_comm, select_device(), nranks, rank, id), "trvb_comm_create");
}

void comm_finalize() {
  if (g_comm != nullptr) { trvb_comm_destroy(g_comm); g_comm = nullptr; }
}

void allreduce(trvb_ctx* ctx, double* buf, long long n) {
  if (g_comm == nullptr) return;
  check(trvb_allreduce(ctx, g_comm, buf, n), "trvb_allreduce");
}

int multi_device_count(const trv::ParameterSet& params) {
  if (g_comm != nullptr || g_prof_enabled()) return 1;
  for (const char* name : {"TRV_GPU_DEVICE", "LOCAL_RANK"}) {
    const char* v = std::getenv(name);
    if (v != nullptr && v[0] != '\0') return 1;
  }
  const char* multi = std::getenv("TRV_GPU_MULTI");
  if (multi != nullptr && multi[0] == '0') return 1;
  if (params.nmesh < (1LL << 24) && !(multi != nullptr && multi[0] == '1')) return 1;
  const int n = trvs::get_gpu_count();
  return n > 1 ? n : 1;
}

void release_contexts() {
  std::lock_guard<std::mutex> lock(g_ctx_mutex);
  g_ctx_recent.clear();
  g_last_ctx = nullptr;
  // With the contexts gone the arena's cached blocks serve nobody: back to the driver.
  const int n = trvb_device_count();
  for (int d = 0; d < n; d++) trvb_arena_release(d);
}

Catalogue::Catalogue(std::shared_ptr<trvb_ctx> ctx, ParticleCatalogue& particles,
                     LineOfSight* los, bool need_weights) : ctx_(ctx) {
  if (particles.pdata == nullptr) {
    throw trvs::InvalidDataError("Particle data are uninitialised.");
  }
  // The reference's AoS record {x, y, z, nz, ws, wc, w} goes over as is and is
  // transposed to SoA on the device; `w` is dropped when only unit weights
  // are needed.
  (void)need_weights;
  check(trvb_cat_create_aos(ctx_.get(), &cat_, particles.ntotal,
                            reinterpret_cast<const double*>(particles.pdata),
                            los ? reinterpret_cast<const double*>(los) : nullptr),
        "trvb_cat_create_aos");
}

Catalogue::Catalogue(std::shared_ptr<trvb_ctx> ctx, long long n, const double* x,
                     const double* y, const double* z, const double* w,
                     const double* los, bool on_device) : ctx_(ctx) {
  // Device arrays are borrowed for the life of this catalogue (one estimator call).
  check(trvb_cat_create(ctx_.get(), &cat_, n, x, y, z, w, los, on_device ? 2 : 0),
        "trvb_cat_create");
}

Catalogue::~Catalogue() { trvb_cat_destroy(cat_); }

namespace {
bool g_prof_on = false;
std::chrono::steady_clock::time_point g_prof_last;
std::vector<std::pair<std::string, double> > g_prof;
}  // namespace

void profile_enable(bool on) { g_prof_on = on; }
bool profile_enabled() { return g_prof_on; }

void profile_reset() {
  g_prof.clear();
  g_prof_last = std::chrono::steady_clock::now();
}

void profile_mark(trvb_ctx* ctx, const char* phase) {
  if (!g_prof_on) return;
  trvb_ctx_sync(ctx);
  auto now = std::chrono::steady_clock::now();
  const double dt = std::chrono::duration<double>(now - g_prof_last).count();
  g_prof_last = now;
  for (auto& kv : g_prof) {
    if (kv.first == phase) { kv.second += dt; return; }
  }
  g_prof.emplace_back(phase, dt);
}

std::string profile_report() {
  std::string out = "{";
  char buf[128];
  for (size_t i = 0; i < g_prof.size(); i++) {
    std::snprintf(buf, sizeof(buf), "%s\"%s\": %.6f", i ? ", " : "",
                  g_prof[i].first.c_str(), g_prof[i].second);
    out += buf;
  }
  return out + "}";
}

void Catalogue::set_custom_weights(const double* weights) {
  check(trvb_cat_set_custom_weights(ctx_.get(), cat_, weights),
        "trvb_cat_set_custom_weights");
}

Mesh::Mesh(std::shared_ptr<trvb_ctx> owner, trvb_ctx* grid, int layout)
  : owner_(owner), layout_(layout) {
  bytes_ = trvb_mesh_bytes(grid, layout);
  check(trvb_malloc(owner_.get(), &data_, bytes_), "trvb_malloc");
  trvs::gbytesMemGPU += double(bytes_) / (1024. * 1024. * 1024.);
  trvs::update_maxmem(true);
}

void Mesh::release() {
  if (data_ != nullptr) {
    trvb_free(owner_.get(), data_);
    data_ = nullptr;
    trvs::gbytesMemGPU -= double(bytes_) / (1024. * 1024. * 1024.);
  }
}

Mesh::~Mesh() { release(); }

Mesh::Mesh(Mesh&& o) noexcept
  : owner_(std::move(o.owner_)), data_(o.data_), layout_(o.layout_), bytes_(o.bytes_) {
  o.data_ = nullptr; o.bytes_ = 0;
}

Mesh& Mesh::operator=(Mesh&& o) noexcept {
  if (this != &o) {
    release();
    owner_ = std::move(o.owner_);
    data_ = o.data_; layout_ = o.layout_; bytes_ = o.bytes_;
    o.data_ = nullptr; o.bytes_ = 0;
  }
  return *this;
}

}  // namespace dev

// =====================================================================
// MeshField
// =====================================================================

MeshField::MeshField(trv::ParameterSet& params, bool plan_ini, const std::string& name) {
  (void)plan_ini;   // cuFFT plans are cached per context and created lazily
  this->params = params;
  this->name = name;
  trvs::logger.reset_level(params.verbose);

  this->ctx_ = dev::acquire_context(this->params);
  this->mesh_ = dev::Mesh(this->ctx_, this->ctx_.get(), TRVB_COMPLEX);
  trvs::count_cgrid += 1; trvs::count_grid += 1; trvs::update_maxcntgrid();
  if (this->params.interlace == "true") {
    this->mesh_s_ = dev::Mesh(this->ctx_, this->ctx_.get(), TRVB_COMPLEX);
    trvs::count_cgrid += 1; trvs::count_grid += 1; trvs::update_maxcntgrid();
  }
  this->reset_density_field();

  for (int ax = 0; ax < 3; ax++) {
    this->dr[ax] = this->params.boxsize[ax] / this->params.ngrid[ax];
    this->dk[ax] = 2. * M_PI / this->params.boxsize[ax];
  }
  this->vol = this->params.volume;
  this->vol_cell = this->vol / double(this->params.nmesh);
}

MeshField::~MeshField() {
  if (this->field != nullptr) { std::free(this->field); this->field = nullptr; }
  trvs::count_cgrid -= 1; trvs::count_grid -= 1;
  if (!this->mesh_s_.empty()) { trvs::count_cgrid -= 1; trvs::count_grid -= 1; }
}

void MeshField::reset_density_field() {
  trvb_ctx* c = this->ctx_.get();
  dev::check(trvb_memset0(c, this->mesh_.data(), trvb_mesh_bytes(c, TRVB_COMPLEX)),
             "trvb_memset0");
  if (!this->mesh_s_.empty()) {
    dev::check(trvb_memset0(c, this->mesh_s_.data(), trvb_mesh_bytes(c, TRVB_COMPLEX)),
               "trvb_memset0");
  }
  this->host_stale_ = true;
}

void MeshField::sync_host() {
  const size_t bytes = trvb_mesh_bytes(this->ctx_.get(), TRVB_COMPLEX);
  if (this->field == nullptr) {
    this->field = reinterpret_cast<double (*)[2]>(std::malloc(bytes));
    if (this->field == nullptr) throw std::bad_alloc();
  }
  if (this->host_stale_) {
    dev::check(trvb_d2h(this->ctx_.get(), this->field, this->mesh_.data(), bytes),
               "trvb_d2h");
    this->host_stale_ = false;
  }
}

void MeshField::sync_device() {
  if (this->field == nullptr) return;
  const size_t bytes = trvb_mesh_bytes(this->ctx_.get(), TRVB_COMPLEX);
  dev::check(trvb_h2d(this->ctx_.get(), this->mesh_.data(), this->field, bytes),
             "trvb_h2d");
  this->host_stale_ = false;
}

const double (&MeshField::operator[](long long gid))[2] {
  this->sync_host();
  return this->field[gid];
}

void MeshField::assign_kind(ParticleCatalogue& particles, LineOfSight* los, int kind,
                            int ell, int m, double scale, bool accumulate) {
  for (int ax = 0; ax < 3; ax++) {
    const double extent = particles.pos_max[ax] - particles.pos_min[ax];
    if (this->params.boxsize[ax] < extent && trvs::currTask == 0) {
      trvs::logger.warn(
        "Box size in dimension %d is smaller than catalogue extents: %.3f < %.3f.",
        ax, this->params.boxsize[ax], extent);
    }
  }
  dev::Catalogue cat(this->ctx_, particles, los, kind != TRVB_W_UNIT);
  const int mode = this->params.deterministic ? 1 : 0;
  dev::check(trvb_assign(this->ctx_.get(), cat.get(), kind, ell, m, scale,
                         /*density_units=*/1, accumulate ? 1 : 0, /*shifted=*/0, mode,
                         this->mesh_.view()), "trvb_assign");
  if (!this->mesh_s_.empty()) {
    dev::check(trvb_assign(this->ctx_.get(), cat.get(), kind, ell, m, scale, 1,
                           accumulate ? 1 : 0, /*shifted=*/1, mode,
                           this->mesh_s_.view()), "trvb_assign (shadow)");
  }
  this->host_stale_ = true;
}

void MeshField::assign_weighted_field_to_mesh(ParticleCatalogue& particles,
                                              double (*weights)[2]) {
  dev::Catalogue cat(this->ctx_, particles, nullptr, false);
  cat.set_custom_weights(reinterpret_cast<const double*>(weights));
  const int mode = this->params.deterministic ? 1 : 0;
  dev::check(trvb_assign(this->ctx_.get(), cat.get(), TRVB_W_CUSTOM, 0, 0, 1., 1, 0, 0,
                         mode, this->mesh_.view()), "trvb_assign");
  if (!this->mesh_s_.empty()) {
    dev::check(trvb_assign(this->ctx_.get(), cat.get(), TRVB_W_CUSTOM, 0, 0, 1., 1, 0, 1,
                           mode, this->mesh_s_.view()), "trvb_assign (shadow)");
  }
  this->host_stale_ = true;
}

void MeshField::compute_unweighted_field(ParticleCatalogue& particles) {
  this->assign_kind(particles, nullptr, TRVB_W_UNIT, 0, 0, 1., false);
}

void MeshField::compute_unweighted_field_fluctuations_insitu(ParticleCatalogue& particles) {
  this->compute_unweighted_field(particles);
  const double nbar = double(particles.ntotal) / this->vol;   // S/field.cpp:1235
  dev::check(trvb_mesh_add_const(this->ctx_.get(), this->mesh_.view(), -nbar),
             "trvb_mesh_add_const");
}

void MeshField::compute_ylm_wgtd_field(
  ParticleCatalogue& particles_data, ParticleCatalogue& particles_rand,
  LineOfSight* los_data, LineOfSight* los_rand, double alpha, int ell, int m
) {
  // data - alpha * rand (S/field.cpp:1246-1323), the random catalogue being
  // accumulated into the same mesh with scale -alpha instead of a second mesh.
  this->assign_kind(particles_data, los_data, TRVB_W_YLM_W, ell, m, 1., false);
  this->assign_kind(particles_rand, los_rand, TRVB_W_YLM_W, ell, m, -alpha, true);
}

void MeshField::compute_ylm_wgtd_field(
  ParticleCatalogue& particles, LineOfSight* los, double alpha, int ell, int m
) {
  this->assign_kind(particles, los, TRVB_W_YLM_W, ell, m, alpha, false);
}

void MeshField::compute_ylm_wgtd_quad_field(
  ParticleCatalogue& particles_data, ParticleCatalogue& particles_rand,
  LineOfSight* los_data, LineOfSight* los_rand, double alpha, int ell, int m
) {
  // data + alpha^2 * rand with conj(y_lm) w^2 weights (S/field.cpp:1364-1447).
  this->assign_kind(particles_data, los_data, TRVB_W_CYLM_W2, ell, m, 1., false);
  this->assign_kind(particles_rand, los_rand, TRVB_W_CYLM_W2, ell, m,
                    std::pow(alpha, 2), true);
}

void MeshField::compute_ylm_wgtd_quad_field(
  ParticleCatalogue& particles, LineOfSight* los, double alpha, int ell, int m
) {
  this->assign_kind(particles, los, TRVB_W_CYLM_W2, ell, m, std::pow(alpha, 2), false);
}

void MeshField::fourier_transform() {
  trvb_ctx* c = this->ctx_.get();
  // field *= vol_cell, then unnormalised forward FFT (S/field.cpp:1503-1557).
  dev::check(trvb_fft_forward(c, this->mesh_.view(), this->mesh_.view(), this->vol_cell),
             "trvb_fft_forward");
  trvs::count_fft += 1;
  if (!this->mesh_s_.empty()) {
    dev::check(trvb_fft_forward(c, this->mesh_s_.view(), this->mesh_s_.view(),
                                this->vol_cell), "trvb_fft_forward (shadow)");
    trvs::count_fft += 1;
    dev::check(trvb_interlace_combine(c, this->mesh_.view(), this->mesh_s_.view()),
               "trvb_interlace_combine");
  }
  this->host_stale_ = true;
}

void MeshField::inv_fourier_transform() {
  trvb_ctx* c = this->ctx_.get();
  // field /= vol, then unnormalised backward FFT (S/field.cpp:1664-1719).
  dev::check(trvb_mesh_axpby(c, this->mesh_.view(), 1. / this->vol, this->mesh_.view(), 0.),
             "trvb_mesh_axpby");
  dev::check(trvb_fft_inverse(c, this->mesh_.view(), this->mesh_.view()),
             "trvb_fft_inverse");
  trvs::count_ifft += 1;
  this->host_stale_ = true;
}

void MeshField::apply_wide_angle_pow_law_kernel() {
  if (trvs::currTask == 0) {
    trvs::logger.debug("Applying wide-angle power-law kernel to '%s'.", this->name.c_str());
  }
  dev::check(trvb_mesh_pow_law(this->ctx_.get(), this->mesh_.view(),
                               this->params.i_wa + this->params.j_wa), "trvb_mesh_pow_law");
  this->host_stale_ = true;
}

void MeshField::apply_assignment_compensation() {
  dev::check(trvb_compensate(this->ctx_.get(), this->mesh_.view()), "trvb_compensate");
  this->host_stale_ = true;
}

void MeshField::inv_fourier_transform_ylm_wgtd_field_band_limited(
  MeshField& field_fourier, int ell, int m,
  double k_lower, double k_upper, double& k_eff, int& nmodes
) {
  trvb_ctx* c = this->ctx_.get();
  const double edges[2] = {k_lower, k_upper};
  long long count = 0; double ksum = 0.;
  dev::check(trvb_shell_stats(c, edges, 1, 0, &count, &ksum), "trvb_shell_stats");
  nmodes = static_cast<int>(count);
  k_eff = ksum / double(nmodes);                          // S/field.cpp:1905
  // F / nmodes is folded into the filter (the transform is linear).
  dev::check(trvb_shell_ifft(c, c, field_fourier.mesh_.view(), ell, m, k_lower, k_upper,
                             1. / double(nmodes), this->mesh_.view()), "trvb_shell_ifft");
  trvs::count_ifft += 1;
  this->host_stale_ = true;
}

void MeshField::inv_fourier_transform_sjl_ylm_wgtd_field(
  MeshField& field_fourier, int ell, int m,
  trvm::SphericalBesselCalculator& sjl, double r
) {
  trvb_ctx* c = this->ctx_.get();
  dev::check(trvb_sjl_table(c, sjl.order, sjl.y.data(), sjl.c.data(),
                            static_cast<int>(sjl.y.size()), sjl.step), "trvb_sjl_table");
  dev::check(trvb_sjl_ifft(c, field_fourier.mesh_.view(), ell, m, r, 1. / this->vol,
                           this->mesh_.view()), "trvb_sjl_ifft");
  trvs::count_ifft += 1;
  this->host_stale_ = true;
}

double MeshField::calc_grid_based_powlaw_norm(ParticleCatalogue& particles, int order) {
  if (order < 2) {
    throw trvs::InvalidParameterError(
      "calc_grid_based_powlaw_norm: order must be 2 (power spectrum) or 3 (bispectrum); "
      "got %d.", order);
  }
  this->assign_kind(particles, nullptr, TRVB_W_W, 0, 0, 1., false);
  double vol_int = 0.;
  dev::check(trvb_mesh_sum_pow(this->ctx_.get(), this->mesh_.view(), order, &vol_int),
             "trvb_mesh_sum_pow");
  vol_int *= this->vol_cell;
  return 1. / vol_int;
}

// =====================================================================
// FieldStats
// =====================================================================

FieldStats::FieldStats(trv::ParameterSet& params, bool plan_ini) {
  (void)plan_ini;
  this->params = params;
  this->ctx_ = dev::acquire_context(this->params);
  this->reset_stats();
}

trv::BinnedVectors FieldStats::record_binned_vectors(
  trv::Binning& binning, const std::string& save_file
) {
  if (binning.space != "config" && binning.space != "fourier") {
    if (trvs::currTask == 0) trvs::logger.error("Invalid binning space: '%s'.", binning.space.c_str());
    throw trvs::InvalidDataError("Invalid binning space: '%s'.", binning.space.c_str());
  }
  const bool fourier = binning.space == "fourier";
  double step[3];
  for (int ax = 0; ax < 3; ax++) {
    step[ax] = fourier ? 2. * M_PI / this->params.boxsize[ax]
                       : this->params.boxsize[ax] / this->params.ngrid[ax];
  }
  // per bin, in row-major cell order; only indices whose vector can reach bin_max are visited
  std::vector<trv::BinnedVectors> per_bin(binning.num_bins);
  std::vector<int> reach[3];
  for (int ax = 0; ax < 3; ax++) {
    const int n = this->params.ngrid[ax];
    for (int i = 0; i < n; i++) {
      const int s = (i < n / 2) ? i : i - n;
      if (std::fabs(s * step[ax]) < binning.bin_max + step[ax]) reach[ax].push_back(i);
    }
  }
  for (int i : reach[0]) {
    const int n0 = this->params.ngrid[0], n1 = this->params.ngrid[1], n2 = this->params.ngrid[2];
    const double vx = ((i < n0 / 2) ? i : i - n0) * step[0];
    for (int j : reach[1]) {
      const double vy = ((j < n1 / 2) ? j : j - n1) * step[1];
      for (int k : reach[2]) {
        const double vz = ((k < n2 / 2) ? k : k - n2) * step[2];
        const double scale = std::sqrt(vx * vx + vy * vy + vz * vz);
        for (int b = 0; b < binning.num_bins; b++) {
          if (binning.bin_edges[b] <= scale && scale < binning.bin_edges[b + 1]) {
            per_bin[b].vecx.push_back(vx); per_bin[b].vecy.push_back(vy); per_bin[b].vecz.push_back(vz);
            break;
          }
        }
      }
    }
  }
  trv::BinnedVectors out;
  out.num_bins = binning.num_bins;
  for (int b = 0; b < binning.num_bins; b++) {
    for (size_t t = 0; t < per_bin[b].vecx.size(); t++) {
      out.indices.push_back(b);
      out.lower_edges.push_back(binning.bin_edges[b]);
      out.upper_edges.push_back(binning.bin_edges[b + 1]);
      out.vecx.push_back(per_bin[b].vecx[t]);
      out.vecy.push_back(per_bin[b].vecy[t]);
      out.vecz.push_back(per_bin[b].vecz[t]);
      out.count++;
    }
  }
  if (!save_file.empty()) {
    std::FILE* fp = std::fopen(save_file.c_str(), "w");
    if (fp == nullptr) throw trvs::IOError("Failed to open file: %s", save_file.c_str());
    trv::io::print_binned_vectors_to_file(fp, this->params, out);
    std::fclose(fp);
    if (trvs::currTask == 0) {
      trvs::logger.info("Check binned-vectors file for reference: %s", save_file.c_str());
    }
  }
  return out;
}

void FieldStats::reset_stats() {
  std::fill(this->nmodes.begin(), this->nmodes.end(), 0);
  std::fill(this->npairs.begin(), this->npairs.end(), 0);
  std::fill(this->k.begin(), this->k.end(), 0.);
  std::fill(this->r.begin(), this->r.end(), 0.);
  std::fill(this->sn.begin(), this->sn.end(), 0.);
  std::fill(this->pk.begin(), this->pk.end(), 0.);
  std::fill(this->xi.begin(), this->xi.end(), 0.);
}

void FieldStats::resize_stats(int num_bins) {
  this->nmodes.resize(num_bins); this->npairs.resize(num_bins);
  this->k.resize(num_bins); this->r.resize(num_bins);
  this->sn.resize(num_bins); this->pk.resize(num_bins); this->xi.resize(num_bins);
}

bool FieldStats::if_fields_compatible(MeshField& field_a, MeshField& field_b) {
  for (int ax = 0; ax < 3; ax++) {
    if (this->params.boxsize[ax] != field_a.params.boxsize[ax]
        || this->params.boxsize[ax] != field_b.params.boxsize[ax]
        || this->params.ngrid[ax] != field_a.params.ngrid[ax]
        || this->params.ngrid[ax] != field_b.params.ngrid[ax]) return false;
  }
  return this->params.nmesh == field_a.params.nmesh
    && this->params.nmesh == field_b.params.nmesh;
}

void FieldStats::compute_ylm_wgtd_2pt_stats_in_fourier(
  MeshField& field_a, MeshField& field_b, std::complex<double> shotnoise_amp,
  int ell, int m, trv::Binning& kbinning
) {
  this->resize_stats(kbinning.num_bins);
  if (!this->if_fields_compatible(field_a, field_b)) {
    throw trvs::InvalidDataError("Input mesh fields have incompatible physical properties.");
  }
  this->reset_stats();
  const int nb = kbinning.num_bins;
  std::vector<long long> nm(nb);
  std::vector<double> pk2(2 * nb), sn2(2 * nb);
  const double S[2] = {shotnoise_amp.real(), shotnoise_amp.imag()};
  dev::check(trvb_twopt_fourier(this->ctx_.get(), field_a.device_view(),
                                field_b.device_view(), S, ell, m, this->interlaced_(),
                                kbinning.bin_edges.data(), kbinning.bin_centres.data(), nb,
                                nm.data(), this->k.data(), pk2.data(), sn2.data()),
             "trvb_twopt_fourier");
  for (int b = 0; b < nb; b++) {
    this->nmodes[b] = static_cast<int>(nm[b]);
    this->pk[b] = std::complex<double>(pk2[2 * b], pk2[2 * b + 1]);
    this->sn[b] = std::complex<double>(sn2[2 * b], sn2[2 * b + 1]);
  }
}

void FieldStats::compute_ylm_wgtd_2pt_stats_in_config(
  MeshField& field_a, MeshField& field_b, std::complex<double> shotnoise_amp,
  int ell, int m, trv::Binning& rbinning
) {
  this->resize_stats(rbinning.num_bins);
  if (!this->if_fields_compatible(field_a, field_b)) {
    throw trvs::InvalidDataError("Input mesh fields have incompatible physical properties.");
  }
  this->reset_stats();
  trvb_ctx* c = this->ctx_.get();
  dev::Mesh xi3d(this->ctx_, c, TRVB_COMPLEX);
  const double S[2] = {shotnoise_amp.real(), shotnoise_amp.imag()};
  dev::check(trvb_shot_xi(c, field_a.device_view(), field_b.device_view(), S,
                          this->interlaced_(), xi3d.view()), "trvb_shot_xi");
  trvs::count_ifft += 1;
  const int nb = rbinning.num_bins;
  std::vector<long long> np(nb);
  std::vector<double> xi2(2 * nb);
  dev::check(trvb_twopt_config_bin(c, xi3d.view(), ell, m, rbinning.bin_edges.data(),
                                   rbinning.bin_centres.data(), nb, np.data(),
                                   this->r.data(), xi2.data()), "trvb_twopt_config_bin");
  for (int b = 0; b < nb; b++) {
    this->npairs[b] = static_cast<int>(np[b]);
    this->xi[b] = std::complex<double>(xi2[2 * b], xi2[2 * b + 1]);
  }
}

void FieldStats::compute_uncoupled_shotnoise_for_3pcf(
  MeshField& field_a, MeshField& field_b, int ell_a, int m_a, int ell_b, int m_b,
  std::complex<double> shotnoise_amp, trv::Binning& rbinning
) {
  this->resize_stats(rbinning.num_bins);
  if (!this->if_fields_compatible(field_a, field_b)) {
    throw trvs::InvalidDataError("Input mesh fields have incompatible physical properties.");
  }
  this->reset_stats();
  trvb_ctx* c = this->ctx_.get();
  dev::Mesh xi(this->ctx_, c, TRVB_COMPLEX);
  const double S[2] = {shotnoise_amp.real(), shotnoise_amp.imag()};
  dev::check(trvb_shot_xi(c, field_a.device_view(), field_b.device_view(), S, this->interlaced_(), xi.view()),
             "trvb_shot_xi");
  trvs::count_ifft += 1;
  const int nb = rbinning.num_bins;
  std::vector<long long> np(nb);
  std::vector<double> xi2(2 * nb);
  const double parity = std::pow(-1, this->params.ell1 + this->params.ell2);
  dev::check(trvb_shot_3pcf_bin(c, xi.view(), ell_a, m_a, ell_b, m_b,
                                rbinning.bin_edges.data(), rbinning.bin_centres.data(), nb,
                                parity, np.data(), this->r.data(), xi2.data()),
             "trvb_shot_3pcf_bin");
  for (int b = 0; b < nb; b++) {
    this->npairs[b] = static_cast<int>(np[b]);
    this->xi[b] = std::complex<double>(xi2[2 * b], xi2[2 * b + 1]);
  }
}

std::complex<double> FieldStats::compute_uncoupled_shotnoise_for_bispec_per_bin(
  MeshField& field_a, MeshField& field_b, int ell_a, int m_a, int ell_b, int m_b,
  trvm::SphericalBesselCalculator& sj_a, trvm::SphericalBesselCalculator& sj_b,
  std::complex<double> shotnoise_amp, double k_a, double k_b
) {
  if (!this->if_fields_compatible(field_a, field_b)) {
    throw trvs::InvalidDataError("Input mesh fields have incompatible physical properties.");
  }
  trvb_ctx* c = this->ctx_.get();
  dev::Mesh xi(this->ctx_, c, TRVB_COMPLEX);
  const double S[2] = {shotnoise_amp.real(), shotnoise_amp.imag()};
  dev::check(trvb_shot_xi(c, field_a.device_view(), field_b.device_view(), S, this->interlaced_(), xi.view()),
             "trvb_shot_xi");
  trvs::count_ifft += 1;
  dev::check(trvb_sjl_table(c, sj_a.order, sj_a.y.data(), sj_a.c.data(),
                            static_cast<int>(sj_a.y.size()), sj_a.step), "trvb_sjl_table");
  dev::check(trvb_sjl_table(c, sj_b.order, sj_b.y.data(), sj_b.c.data(),
                            static_cast<int>(sj_b.y.size()), sj_b.step), "trvb_sjl_table");
  double out[2];
  dev::check(trvb_shot_bispec_reduce(c, xi.view(), ell_a, m_a, ell_b, m_b, &k_a, &k_b, 1,
                                     out), "trvb_shot_bispec_reduce");
  return std::complex<double>(out[0], out[1]);
}

}  // namespace trv
