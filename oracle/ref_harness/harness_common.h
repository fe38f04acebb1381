// harness_common.h -- shared helpers of the reference harnesses (test infrastructure).
//
// Each harness is a tiny main() that #includes ONE untouched case header from
// /root/reference, builds the LBM_* functor exactly as the reference driver does,
// runs the reference's own hot line
//     for_each(execution::par_unseq, lattice, lattice + nelem, lbm); *parity = 1 - *parity;
// and dumps binary fp64 populations / macroscopic fields (the reference's own VTK is
// 6-digit ASCII, useless at 1e-10) or prints MLUPS for the CPU baseline.
// No reference source is copied: the headers are compiled where they lie.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <execution>
#include <map>
#include <string>
#include <thread>
#include <vector>

struct Args {
    std::map<std::string, std::string> kv;
    Args(int argc, char** argv) {
        for (int i = 1; i < argc; ++i) {
            std::string s(argv[i]);
            auto p = s.find('=');
            if (p == std::string::npos) { std::fprintf(stderr, "bad arg %s (want key=value)\n", argv[i]); std::exit(2); }
            kv[s.substr(0, p)] = s.substr(p + 1);
        }
    }
    double d(const char* k, double def) const { auto it = kv.find(k); return it == kv.end() ? def : std::stod(it->second); }
    int i(const char* k, int def) const { auto it = kv.find(k); return it == kv.end() ? def : std::stoi(it->second); }
    std::string s(const char* k, const char* def) const { auto it = kv.find(k); return it == kv.end() ? def : it->second; }
};

// The reference's hot line, `steps` times.  threads<=1: the stock call (libstdc++ PSTL;
// serial backend when TBB is absent).  threads>1: the same functor over the same index
// range sharded across std::threads -- identical arithmetic, race-free by construction
// (two-lattice push, SURVEY.md 5 "race detection").
template <class LBM>
double run_steps(LBM& lbm, double* lattice, size_t nelem, int* parity, int steps, int threads)
{
    auto t0 = std::chrono::high_resolution_clock::now();
    for (int s = 0; s < steps; ++s) {
        if (threads <= 1) {
            std::for_each(std::execution::par_unseq, lattice, lattice + nelem, lbm);
        } else {
            std::vector<std::thread> pool;
            for (int t = 0; t < threads; ++t) {
                size_t b = nelem * t / threads, e = nelem * (t + 1) / threads;
                pool.emplace_back([&lbm, lattice, b, e]() { std::for_each(lattice + b, lattice + e, lbm); });
            }
            for (auto& th : pool) th.join();
        }
        *parity = 1 - *parity;
    }
    auto t1 = std::chrono::high_resolution_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

struct Dump {
    FILE* f;
    explicit Dump(const std::string& path) : f(path.empty() ? nullptr : std::fopen(path.c_str(), "wb")) {}
    ~Dump() { if (f) std::fclose(f); }
    void put(const double* p, size_t n) { if (f) std::fwrite(p, sizeof(double), n, f); }
    void put(const std::vector<double>& v) { put(v.data(), v.size()); }
    void put_u8(const uint8_t* p, size_t n) { if (f) std::fwrite(p, 1, n, f); }
};

inline void report(const char* name, size_t nelem, int steps, int threads, double sec)
{
    std::printf("{\"case\": \"%s\", \"nelem\": %zu, \"steps\": %d, \"threads\": %d, \"seconds\": %.6f, \"mlups\": %.6f}\n",
                name, nelem, steps, threads, sec, sec > 0 ? nelem * (double)steps / sec / 1e6 : 0.0);
}
