// case_common.h -- shared plumbing of the B200 case drivers: the key/value config reader (with the
// reference's reading order), the device lattice handle (the counterpart of an LBM_* functor object),
// legacy-VTK output and the MLUPS report.
//
// Surface mirrored from the reference case headers:
//   config   : `key value # comment` lines, first line skipped by the older drivers' reader
//              (SC/apps/laplace2D.h:417-437, PF/apps/rayleighTaylor2D.h:877-902)
//   derived  : nu = ulb N / Re, omega = 1/(3 nu + 0.5), dx = 1/N, dt = dx ulb  (SC/apps/laplace2D.h:52-58)
//   output   : sol_%07d.vtk legacy ASCII STRUCTURED_POINTS, y-outer/x-inner (SC/apps/laplace2D.h:319-365);
//              "result: <s> seconds / <MLUPS> MLUPS" (SC/apps/laplace2D.h:79-86)
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/clbm.h"

namespace coolbm {

// Reads a config file the way the older reference drivers do: each getline() result is discarded and the next
// two whitespace tokens of the stream are (param, value); hence line 1 never contributes and only the first
// two tokens of every later line count.  Unknown keys are reported through `unknown`.
inline std::map<std::string, std::string> read_config(const std::string &path, const std::string &case_name)
{
    std::ifstream in(path);
    if (!in.is_open())
        throw std::invalid_argument("Config file not found. It should be named \"" + case_name + "\" in Files_Config.");
    std::map<std::string, std::string> kv;
    std::string line, param, value;
    while (std::getline(in, line)) {
        if (!(in >> param >> value)) break;
        kv[param] = value;
    }
    return kv;
}

struct Config {
    std::map<std::string, std::string> kv;
    std::map<std::string, bool> used;
    double d(const char *k, double def) { used[k] = true; auto it = kv.find(k); return it == kv.end() ? def : std::stod(it->second); }
    int i(const char *k, int def) { used[k] = true; auto it = kv.find(k); return it == kv.end() ? def : std::stoi(it->second); }
    bool has(const char *k) const { return kv.count(k) != 0; }
    void warn_unknown() const
    {
        for (auto &p : kv)
            if (!used.count(p.first)) std::cerr << "Warning: unknown parameter \"" << p.first << "\"\n";
    }
};

// Optional keys of this library, absent from the reference's config files: `collision MRT` selects the moment-space
// collision operator (include/clbm.h, CLBM_COLLISION_MRT: every Shan-Chen and HCZ model), `s_e`, `s_eps`, `s_q` are
// its free rates (default: omega, which is the BGK operator evaluated in moment space).
inline void apply_collision_keys(Config &cfg, clbm_params &p, double omega)
{
    if (!cfg.has("collision")) return;
    cfg.used["collision"] = true;
    const std::string v = cfg.kv["collision"];
    if (!(v == "MRT" || v == "mrt" || v == "1")) return;
    const double s_e = cfg.d("s_e", -1), s_eps = cfg.d("s_eps", -1), s_q = cfg.d("s_q", -1);
    p.collision = CLBM_COLLISION_MRT;
    p.s_e = s_e > 0 ? s_e : omega; p.s_eps = s_eps > 0 ? s_eps : omega; p.s_q = s_q > 0 ? s_q : omega;
    std::cout << "collision = MRT  s_e = " << p.s_e << "  s_eps = " << p.s_eps << "  s_q = " << p.s_q << "  (s_nu = omega)\n";
}

struct LbParameters { double nu, omega, dx, dt; };
inline LbParameters lb_parameters(double ulb, int lref, double Re)
{
    LbParameters p;
    p.nu = ulb * lref / Re;
    p.omega = 1. / (3. * p.nu + 0.5);
    p.dx = 1. / lref;
    p.dt = p.dx * ulb;
    return p;
}

// clbm_params with the library defaults (single slab on the current device, fused kernels)
inline clbm_params default_params(int model, int nx, int ny, int nz)
{
    clbm_params p{};
    p.abi_version = CLBM_ABI_VERSION;
    p.model = model;
    p.nx = nx; p.ny = ny; p.nz = nz;
    p.nx_global = nx; p.x_offset = 0;
    p.device = -1;
    p.fused = 1;
    p.omega = 1.0;
    p.a = 1.0; p.b = 4.0; p.R = 1.0;
    return p;
}

inline void check(int rc)
{
    if (rc != CLBM_OK) throw std::runtime_error(std::string("clbm: ") + clbm_last_error());
}

// Device-resident lattice: what `LBM_* lbm{lattice, flag, parity, ...}` is in the reference drivers.
class DeviceLattice {
public:
    clbm_params prm{};
    clbm_ctx *ctx = nullptr;
    explicit DeviceLattice(const clbm_params &p) : prm(p) { check(clbm_create(&prm, &ctx)); }
    ~DeviceLattice() { clbm_destroy(ctx); }
    DeviceLattice(const DeviceLattice &) = delete;
    DeviceLattice &operator=(const DeviceLattice &) = delete;

    size_t nelem() const { return (size_t)prm.nx * prm.ny * prm.nz; }
    void init_case(int id, std::vector<double> args) { check(clbm_init_case(ctx, id, args.data(), (int)args.size())); }
    void step(int n = 1) { check(clbm_step(ctx, n)); }
    void sync() { check(clbm_sync(ctx)); }
    double reduce(int kind) { double v = 0; check(clbm_reduce(ctx, kind, &v)); return v; }
    struct Fields { std::vector<double> s0, s1, s2, ux, uy, uz; std::vector<uint8_t> flag; };
    Fields fields(bool want_s1 = true, bool want_u = true)
    {
        Fields f;
        const size_t n = nelem();
        f.s0.resize(n); f.s2.resize(n); f.flag.resize(n);
        if (want_s1) f.s1.resize(n);
        if (want_u) { f.ux.resize(n); f.uy.resize(n); f.uz.resize(n); }
        check(clbm_download_fields(ctx, f.s0.data(), want_s1 ? f.s1.data() : nullptr, f.s2.data(), want_u ? f.ux.data() : nullptr,
                                   want_u ? f.uy.data() : nullptr, want_u ? f.uz.data() : nullptr, f.flag.data()));
        return f;
    }
};

// ---- VTK output -----------------------------------------------------------------------------------
// Default: the reference's legacy ASCII STRUCTURED_POINTS file, token for token (sol_%07d.vtk).  The ASCII writer
// dominates the wall time of a large case, so the same calls can instead produce an XML ImageData file
// with raw appended binary blocks (sol_%07d.vti, Float32 like the reference's `float` columns, little endian, UInt64 block
// headers): set COOLBM_VTK_FORMAT=vti in the environment.  Point order in the .vti is VTK's own (x fastest, then y, then z
// ascending); the legacy file keeps the reference's loop order (z descending in 3-D).
class VtkWriter {
    std::ofstream os;
    int nx, ny, nz;
    bool binary;
    std::string header;                 // <DataArray .../> lines of the .vti
    std::vector<char> appended;         // [UInt64 nbytes][payload] blocks
    double dx_;

    static float to_float(double v) { return (float)v; }
    static float to_float(float v) { return v; }
    static float to_float(int v) { return (float)v; }
    static float to_float(const char *v) { return (float)std::atof(v); }
    static float to_float(const std::string &v) { return (float)std::atof(v.c_str()); }
    void begin_block(const char *name, const char *type, int ncomp, size_t nvalues, size_t elem)
    {
        std::ostringstream h;
        h << "        <DataArray type=\"" << type << "\" Name=\"" << name << "\"";
        if (ncomp > 1) h << " NumberOfComponents=\"" << ncomp << "\"";
        h << " format=\"appended\" offset=\"" << appended.size() << "\"/>\n";
        header += h.str();
        const uint64_t nbytes = (uint64_t)nvalues * elem;
        const char *p = reinterpret_cast<const char *>(&nbytes);
        appended.insert(appended.end(), p, p + sizeof(nbytes));
    }
    template <class T> void push(T v)
    {
        const char *p = reinterpret_cast<const char *>(&v);
        appended.insert(appended.end(), p, p + sizeof(T));
    }
public:
    static bool binary_requested()
    {
        const char *e = std::getenv("COOLBM_VTK_FORMAT");
        return e && (std::string(e) == "vti" || std::string(e) == "binary");
    }
    VtkWriter(int time_iter, int nx_, int ny_, int nz_, double dx) : nx(nx_), ny(ny_), nz(nz_), binary(binary_requested()), dx_(dx)
    {
        std::stringstream ss;
        ss << "sol_" << std::setfill('0') << std::setw(7) << time_iter << (binary ? ".vti" : ".vtk");
        os.open(ss.str(), binary ? std::ios::binary : std::ios::out);
        if (binary) return;
        os << "# vtk DataFile Version 2.0\n" << "iteration " << time_iter << "\nASCII\n\n";
        os << "DATASET STRUCTURED_POINTS\n" << "DIMENSIONS " << nx << " " << ny << " " << nz << "\n";
        os << "ORIGIN 0 0 0\n" << "SPACING " << dx << " " << dx << " " << dx << "\n\n";
        os << "POINT_DATA " << (size_t)nx * ny * nz << "\n";
    }
    ~VtkWriter()
    {
        if (!binary) return;
        os << "<?xml version=\"1.0\"?>\n"
           << "<VTKFile type=\"ImageData\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n"
           << "  <ImageData WholeExtent=\"0 " << nx - 1 << " 0 " << ny - 1 << " 0 " << nz - 1 << "\" Origin=\"0 0 0\" Spacing=\""
           << std::setprecision(17) << dx_ << " " << dx_ << " " << dx_ << "\">\n"
           << "    <Piece Extent=\"0 " << nx - 1 << " 0 " << ny - 1 << " 0 " << nz - 1 << "\">\n      <PointData>\n"
           << header << "      </PointData>\n    </Piece>\n  </ImageData>\n  <AppendedData encoding=\"raw\">\n_";
        os.write(appended.data(), (std::streamsize)appended.size());
        os << "\n  </AppendedData>\n</VTKFile>\n";
    }
    // value(i) with i = z + nz*(y + ny*x); loop order of the reference writers: z descending (3-D), y outer, x inner
    // plane_blank: a blank line after every z plane even in 2-D (the PF writers' Flag block, PF/apps/rayleighTaylor2D.h:763-780)
    template <class V> void scalars(const char *name, const char *type, V value, bool plane_blank = false)
    {
        if (binary) {
            const bool is_int = std::string(type) == "int";
            begin_block(name, is_int ? "Int32" : "Float32", 1, (size_t)nx * ny * nz, 4);
            for (int z = 0; z < nz; ++z)
                for (int y = 0; y < ny; ++y)
                    for (int x = 0; x < nx; ++x) {
                        const float v = to_float(value((size_t)z + (size_t)nz * (y + (size_t)ny * x)));
                        if (is_int) push<int32_t>((int32_t)v); else push<float>(v);
                    }
            return;
        }
        os << "SCALARS " << name << " " << type << " 1\nLOOKUP_TABLE default\n";
        for (int z = nz - 1; z >= 0; --z) {
            for (int y = 0; y < ny; ++y) {
                for (int x = 0; x < nx; ++x) os << value((size_t)z + (size_t)nz * (y + (size_t)ny * x)) << " ";
                os << "\n";
            }
            if (nz > 1 || plane_blank) os << "\n";
        }
        os << "\n";
    }
    template <class V> void vectors_binary(const char *name, V value, bool third_is_zero)
    {
        begin_block(name, "Float32", 3, (size_t)3 * nx * ny * nz, 4);
        for (int z = 0; z < nz; ++z)
            for (int y = 0; y < ny; ++y)
                for (int x = 0; x < nx; ++x) {
                    auto v = value((size_t)z + (size_t)nz * (y + (size_t)ny * x));
                    push<float>((float)v[0]);
                    push<float>((float)v[1]);
                    push<float>(third_is_zero ? 0.0f : (float)v[2]);
                }
    }
    // "u v 0" per node with a blank line after every row (the AB writers, AB/apps/Young_Laplace2D.h:406-412)
    template <class V> void vectors_rows(const char *name, V value)
    {
        if (binary) { vectors_binary(name, value, true); return; }
        os << "VECTORS " << name << " float\n";
        for (int y = 0; y < ny; ++y) {
            for (int x = 0; x < nx; ++x) {
                auto v = value((size_t)y + (size_t)ny * x);
                os << v[0] << " " << v[1] << " 0\n";
            }
            os << "\n";
        }
        os << "\n";
    }
    template <class V> void vectors(const char *name, V value)
    {
        if (binary) { vectors_binary(name, value, false); return; }
        os << "VECTORS " << name << " float\n";
        for (int z = nz - 1; z >= 0; --z)
            for (int y = 0; y < ny; ++y)
                for (int x = 0; x < nx; ++x) {
                    auto v = value((size_t)z + (size_t)nz * (y + (size_t)ny * x));
                    os << v[0] << " " << v[1] << " " << v[2] << "\n";
                }
        os << "\n";
    }
};

// ---- clock / MLUPS ----------------------------------------------------------------------------------
struct Stopwatch {
    std::chrono::high_resolution_clock::time_point t0 = std::chrono::high_resolution_clock::now();
    int iters = 0;
    // "result: <s> seconds / result: <v> MLUPS" (SC/apps/laplace2D.h:79-86); the PF / AB drivers print
    // "Runtime: ... / Throughput: ..." (PF/apps/laplace3D.h:101-112, AB/apps/PulsatileBloodFlow2D.h:711-717)
    void report(size_t nelem, const char *l1 = "result: ", const char *l2 = "result: ", const char *unit1 = " seconds\n") const
    {
        auto us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::high_resolution_clock::now() - t0).count();
        double mlups = (double)(nelem * (size_t)iters) / (double)us;
        std::cout << l1 << us / 1e6 << unit1;
        std::cout << l2 << std::setprecision(4) << mlups << " MLUPS" << std::endl;
    }
};

inline void progress_line(int time_iter, double dt, double max_t, bool flush = false)
{
    std::cout << "Saving profiles at iteration " << time_iter << ", t = " << std::setprecision(4) << time_iter * dt
              << std::setprecision(3) << " [" << time_iter * dt / max_t * 100. << "%]\n";
    if (flush) std::cout.flush();
}

// runs `steps_total` device steps, stopping at every multiple of the output frequencies for the callback
inline void run_loop(DeviceLattice &lat, int steps_total, int out_freq, int vtk_freq, Stopwatch &sw,
                     const std::function<void(int, bool, bool)> &on_output)
{
    int t = 0;
    while (t < steps_total) {
        const bool v = vtk_freq != 0 && t % vtk_freq == 0, o = out_freq != 0 && t % out_freq == 0;
        if (v || o) on_output(t, v, o);
        int next = steps_total;
        if (vtk_freq != 0) next = std::min(next, (t / vtk_freq + 1) * vtk_freq);
        if (out_freq != 0) next = std::min(next, (t / out_freq + 1) * out_freq);
        lat.step(next - t);           // the reference's hot line, (next - t) times, without leaving the device
        sw.iters += next - t;
        t = next;
    }
    lat.sync();
}

}  // namespace coolbm
