// hermespy_rt_module.cpp -- Python module `hermespy_rt`, the consumer-facing
// surface of the reference (compute_paths_pybind11.cpp:99-210), rebuilt over
// the B200 library.  Same module name, function signature, keyword names and
// ChannelInfo attributes (num_paths, directions_rx, directions_tx, a_te, a_tm,
// tau, freq_shift; shapes as asserted by the reference's test/test.py:61-87).
//
// Differences from the reference binding, all of them fixes it needs anyway
// (SURVEY section 8 row f1): the C headers are wrapped in extern "C" (the
// reference module does not import on Linux); inputs are converted to
// C-contiguous float32 and must have shape (n, 3); outputs are zero-initialised
// numpy arrays owned by Python (no leaks, no uninitialised slots); the complex
// gains are written by the device straight into the complex64 arrays Python
// receives (compute_paths_c64: no re / im repack loop, :22-42); the scene file is
// parsed once per (path, size, mtime) and the uploaded scene + BVH are reused
// across calls (the reference re-reads the file on every call, :119); RaysInfo
// is not materialised (the reference computes and discards it, :172-176); the
// GIL is released while the GPU works.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <sys/stat.h>

#include <complex>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/hrt_cuda.h"   // extern "C" inside

namespace py = pybind11;
using farr = py::array_t<float, py::array::c_style | py::array::forcecast>;

namespace {

struct Channel {
  size_t num_paths;
  py::array_t<float> directions_rx, directions_tx, tau, freq_shift;
  py::array_t<std::complex<float>> a_te, a_tm;
};

const Vec3 *as_vec3(const farr &a, size_t n, const char *what)
{
  if (a.ndim() != 2 || a.shape(0) != (py::ssize_t)n || a.shape(1) != 3)
    throw std::invalid_argument(std::string(what) + ": expected an array of shape (" + std::to_string(n) + ", 3)");
  return reinterpret_cast<const Vec3 *>(a.data());
}

// The scene of the last call, kept parsed while the file is unchanged; the
// library in turn keeps the uploaded copy and its BVH while the content is the
// same (csrc/compute_paths.c).  One call at a time uses it (the implicit
// context of compute_paths() is not re-entrant, like the reference).
struct SceneCache {
  std::mutex mu;
  std::string path; off_t size = -1; time_t mtime_s = 0; long mtime_ns = 0;
  Scene scene{}; bool loaded = false;
  Scene *get(const std::string &p)
  {
    struct stat st;
    if (stat(p.c_str(), &st) != 0) throw std::runtime_error("hermespy_rt: cannot open scene file " + p);
    if (loaded && p == path && st.st_size == size && st.st_mtim.tv_sec == mtime_s && st.st_mtim.tv_nsec == mtime_ns) return &scene;
    if (loaded) { free_scene(&scene); loaded = false; }
    scene = scene_load(p.c_str());
    path = p; size = st.st_size; mtime_s = st.st_mtim.tv_sec; mtime_ns = st.st_mtim.tv_nsec; loaded = true;
    return &scene;
  }
};
SceneCache g_scene_cache;

py::array_t<float> zeros(std::vector<py::ssize_t> shape)
{
  py::array_t<float> a(shape);
  std::fill_n(a.mutable_data(), a.size(), 0.f);
  return a;
}

Channel make_channel(size_t R, size_t T, size_t n)
{
  Channel c;
  c.num_paths = n;
  const auto r = (py::ssize_t)R, t = (py::ssize_t)T, k = (py::ssize_t)n;
  c.directions_rx = zeros({r, t, k, 3});
  c.directions_tx = zeros({r, t, k, 3});
  c.tau = zeros({r, t, k});
  c.freq_shift = zeros({r, t, k});
  return c;
}

py::array_t<std::complex<float>> czeros(size_t R, size_t T, size_t n)
{
  py::array_t<std::complex<float>> a({(py::ssize_t)R, (py::ssize_t)T, (py::ssize_t)n});
  std::fill_n(a.mutable_data(), a.size(), std::complex<float>(0.f, 0.f));
  return a;
}

std::pair<Channel, Channel> compute_paths_py(const std::string &mesh_filepath, farr rx_positions,
                                             farr tx_positions, farr rx_velocities, farr tx_velocities,
                                             float carrier_frequency, size_t num_rx, size_t num_tx,
                                             size_t num_paths, size_t num_bounces)
{
  if (!num_rx || !num_tx || !num_paths || !num_bounces || !(carrier_frequency > 0.f))
    throw std::invalid_argument("num_rx, num_tx, num_paths, num_bounces and carrier_frequency must be > 0");
  const Vec3 *rx = as_vec3(rx_positions, num_rx, "rx_positions");
  const Vec3 *tx = as_vec3(tx_positions, num_tx, "tx_positions");
  const Vec3 *rxv = as_vec3(rx_velocities, num_rx, "rx_velocities");
  const Vec3 *txv = as_vec3(tx_velocities, num_tx, "tx_velocities");
  if (hrt_device_count() <= 0)
    throw std::runtime_error("hermespy_rt: no CUDA device available (this build has no CPU path)");

  const size_t nl = num_rx * num_tx;
  Channel los = make_channel(num_rx, num_tx, 1), sc = make_channel(num_rx, num_tx, num_bounces * num_paths);
  los.a_te = czeros(num_rx, num_tx, 1); los.a_tm = czeros(num_rx, num_tx, 1);
  sc.a_te = czeros(num_rx, num_tx, num_bounces * num_paths); sc.a_tm = czeros(num_rx, num_tx, num_bounces * num_paths);
  std::vector<float> l_re[2], l_im[2];
  for (int k = 0; k < 2; ++k) { l_re[k].assign(nl, 0.f); l_im[k].assign(nl, 0.f); }

  ChannelInfo ci_los = { 1, (Vec3 *)los.directions_rx.mutable_data(), (Vec3 *)los.directions_tx.mutable_data(),
                         l_re[0].data(), l_im[0].data(), l_re[1].data(), l_im[1].data(),
                         los.tau.mutable_data(), los.freq_shift.mutable_data() };
  ChannelInfo ci_sc = { (uint32_t)(num_bounces * num_paths), (Vec3 *)sc.directions_rx.mutable_data(),
                        (Vec3 *)sc.directions_tx.mutable_data(), nullptr, nullptr, nullptr, nullptr,
                        sc.tau.mutable_data(), sc.freq_shift.mutable_data() };
  float *te = reinterpret_cast<float *>(sc.a_te.mutable_data()), *tm = reinterpret_cast<float *>(sc.a_tm.mutable_data());
  {
    py::gil_scoped_release nogil;
    std::lock_guard<std::mutex> lk(g_scene_cache.mu);
    Scene *scene = g_scene_cache.get(mesh_filepath);
    compute_paths_c64(scene, (Vec3 *)rx, (Vec3 *)tx, (Vec3 *)rxv, (Vec3 *)txv, carrier_frequency,
                      num_rx, num_tx, num_paths, num_bounces, &ci_los, nullptr, &ci_sc, nullptr, te, tm);
  }
  for (size_t k = 0; k < nl; ++k) {                        // num_rx * num_tx line-of-sight values
    los.a_te.mutable_data()[k] = std::complex<float>(l_re[0][k], l_im[0][k]);
    los.a_tm.mutable_data()[k] = std::complex<float>(l_re[1][k], l_im[1][k]);
  }
  return {std::move(los), std::move(sc)};
}

// Extension: impulse response per (rx, tx) of the same path set, reduced on the
// GPU (include/hermespy_rt.h, compute_cir).  Returns (cir, dropped) with cir a
// complex64 array of shape (num_rx, num_tx, num_bins, 2): [..., 0] = TE, [..., 1] = TM.
std::pair<py::array_t<std::complex<float>>, size_t>
compute_cir_py(const std::string &mesh_filepath, farr rx_positions, farr tx_positions, farr rx_velocities,
               farr tx_velocities, float carrier_frequency, size_t num_rx, size_t num_tx, size_t num_paths,
               size_t num_bounces, float tau0, float dt, size_t num_bins)
{
  if (!num_rx || !num_tx || !num_paths || !num_bounces || !num_bins || !(carrier_frequency > 0.f) || !(dt > 0.f))
    throw std::invalid_argument("num_rx, num_tx, num_paths, num_bounces, num_bins, carrier_frequency and dt must be > 0");
  const Vec3 *rx = as_vec3(rx_positions, num_rx, "rx_positions");
  const Vec3 *tx = as_vec3(tx_positions, num_tx, "tx_positions");
  const Vec3 *rxv = as_vec3(rx_velocities, num_rx, "rx_velocities");
  const Vec3 *txv = as_vec3(tx_velocities, num_tx, "tx_velocities");
  if (hrt_device_count() <= 0)
    throw std::runtime_error("hermespy_rt: no CUDA device available (this build has no CPU path)");
  py::array_t<std::complex<float>> cir({(py::ssize_t)num_rx, (py::ssize_t)num_tx, (py::ssize_t)num_bins, (py::ssize_t)2});
  size_t dropped = 0;
  {
    py::gil_scoped_release nogil;
    std::lock_guard<std::mutex> lk(g_scene_cache.mu);
    Scene *scene = g_scene_cache.get(mesh_filepath);
    dropped = compute_cir(scene, (Vec3 *)rx, (Vec3 *)tx, (Vec3 *)rxv, (Vec3 *)txv, carrier_frequency, num_rx, num_tx,
                          num_paths, num_bounces, tau0, dt, num_bins, reinterpret_cast<float *>(cir.mutable_data()));
  }
  return {std::move(cir), dropped};
}

// Extension: the valid scatter paths as a structured array (no dense slots for
// dead rays / occluded receivers).  Returns (records[:min(found, capacity)], found).
std::pair<py::array, size_t>
compute_path_list_py(const std::string &mesh_filepath, farr rx_positions, farr tx_positions, farr rx_velocities,
                     farr tx_velocities, float carrier_frequency, size_t num_rx, size_t num_tx, size_t num_paths,
                     size_t num_bounces, size_t capacity)
{
  if (!num_rx || !num_tx || !num_paths || !num_bounces || !capacity || !(carrier_frequency > 0.f))
    throw std::invalid_argument("num_rx, num_tx, num_paths, num_bounces, capacity and carrier_frequency must be > 0");
  const Vec3 *rx = as_vec3(rx_positions, num_rx, "rx_positions");
  const Vec3 *tx = as_vec3(tx_positions, num_tx, "tx_positions");
  const Vec3 *rxv = as_vec3(rx_velocities, num_rx, "rx_velocities");
  const Vec3 *txv = as_vec3(tx_velocities, num_tx, "tx_velocities");
  if (hrt_device_count() <= 0)
    throw std::runtime_error("hermespy_rt: no CUDA device available (this build has no CPU path)");
  std::vector<HrtPathRecord> buf(capacity);
  size_t found = 0;
  {
    py::gil_scoped_release nogil;
    std::lock_guard<std::mutex> lk(g_scene_cache.mu);
    Scene *scene = g_scene_cache.get(mesh_filepath);
    found = compute_path_list(scene, (Vec3 *)rx, (Vec3 *)tx, (Vec3 *)rxv, (Vec3 *)txv, carrier_frequency, num_rx, num_tx,
                              num_paths, num_bounces, buf.data(), capacity);
  }
  const size_t kept = found < capacity ? found : capacity;
  py::list fields;
  const char *names[] = {"path", "rx", "tx", "bounce", "a_te_re", "a_te_im", "a_tm_re", "a_tm_im", "tau", "freq_shift", "direction_rx"};
  const char *fmts[] = {"<u4", "<u4", "<u2", "<u2", "<f4", "<f4", "<f4", "<f4", "<f4", "<f4", "<f4"};
  for (int k = 0; k < 11; ++k) {
    if (k == 10) fields.append(py::make_tuple(names[k], fmts[k], py::make_tuple(3)));
    else fields.append(py::make_tuple(names[k], fmts[k]));
  }
  py::dtype dt = py::dtype::from_args(fields);
  if ((size_t)dt.itemsize() != sizeof(HrtPathRecord)) throw std::runtime_error("hermespy_rt: path record layout mismatch");
  py::array out(dt, std::vector<py::ssize_t>{(py::ssize_t)kept});
  memcpy(out.mutable_data(), buf.data(), kept * sizeof(HrtPathRecord));
  return {std::move(out), found};
}

}  // namespace

PYBIND11_MODULE(hermespy_rt, m)
{
  m.doc() = "hermespy-rt compute_paths() on NVIDIA B200 (drop-in for the reference module)";
  py::class_<Channel>(m, "ChannelInfo")
      .def_readonly("num_paths", &Channel::num_paths)
      .def_readonly("directions_rx", &Channel::directions_rx)
      .def_readonly("directions_tx", &Channel::directions_tx)
      .def_readonly("a_te", &Channel::a_te)
      .def_readonly("a_tm", &Channel::a_tm)
      .def_readonly("tau", &Channel::tau)
      .def_readonly("freq_shift", &Channel::freq_shift);
  m.def("compute_paths", &compute_paths_py,
        "Gains, delays, angles and Doppler shifts of LoS and scatter paths; returns (los, scatter)",
        py::arg("mesh_filepath"), py::arg("rx_positions"), py::arg("tx_positions"),
        py::arg("rx_velocities"), py::arg("tx_velocities"), py::arg("carrier_frequency"),
        py::arg("num_rx"), py::arg("num_tx"), py::arg("num_paths"), py::arg("num_bounces"));
  m.def("compute_cir", &compute_cir_py,
        "Channel impulse response per (rx, tx) of the compute_paths() path set, reduced on the GPU; "
        "returns (cir[num_rx, num_tx, num_bins, 2] complex64 (TE, TM), number of paths outside the window)",
        py::arg("mesh_filepath"), py::arg("rx_positions"), py::arg("tx_positions"),
        py::arg("rx_velocities"), py::arg("tx_velocities"), py::arg("carrier_frequency"),
        py::arg("num_rx"), py::arg("num_tx"), py::arg("num_paths"), py::arg("num_bounces"),
        py::arg("tau0"), py::arg("dt"), py::arg("num_bins"));
  m.def("compute_path_list", &compute_path_list_py,
        "The valid scatter paths of the compute_paths() path set as a structured array of 48-byte records "
        "(path, rx, tx, bounce, a_te_re, a_te_im, a_tm_re, a_tm_im, tau, freq_shift, direction_rx); "
        "returns (records, number of valid paths found)",
        py::arg("mesh_filepath"), py::arg("rx_positions"), py::arg("tx_positions"),
        py::arg("rx_velocities"), py::arg("tx_velocities"), py::arg("carrier_frequency"),
        py::arg("num_rx"), py::arg("num_tx"), py::arg("num_paths"), py::arg("num_bounces"), py::arg("capacity"));
}
