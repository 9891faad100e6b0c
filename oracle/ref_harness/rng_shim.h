/*
 * rng_shim.h -- pre-included (-include) into the reference translation units that draw random
 * numbers (src/scene/light.cpp, src/rendering/render_utils.cpp).  It swaps
 *     std::random_device, std::mt19937, std::uniform_int_distribution
 * for harness types that pull every draw from the counter-based stream of include/romis_rng.h
 * (north_star: "both sides are driven by the same counter-based RNG, injected into the reference
 * through a test harness").  rand() (reservoir.cpp:24, light.cpp:20,28-29), powf and expf are
 * replaced at link time in shims.cpp.  The reference sources themselves are compiled unmodified.
 */
#pragma once
// Pull in every standard header that mentions the names we are about to re-route, so that the
// library's own code is parsed before the macros exist.
#include <algorithm>
#include <array>
#include <chrono>
#include <filesystem>
#include <format>
#include <functional>
#include <iostream>
#include <memory>
#include <numeric>
#include <optional>
#include <random>
#include <span>
#include <string>
#include <unordered_map>
#include <variant>
#include <vector>
#include <cstdint>

extern "C" uint32_t romis_shim_engine_next(void);   // next ENGINE-stream draw for the current pixel
extern "C" void romis_shim_engine_ctor(void);       // a std::mt19937 was constructed
extern "C" int romis_shim_asis(void);               // "as-is" timing arm: the reference's own random sources, at their own cost

namespace romis_shim {
// As-is mode (bench.py's second cpu_baseline entry): the real std::random_device is constructed and drawn
// and a real std::mt19937 is seeded and drawn wherever the reference does so (once per PIXEL in genCanonicalSamples,
// light.cpp:49-50), and rand() is glibc's (shims.cpp) -- the costs the reference pays.  These lines are parsed before the
// macros below exist, so std::random_device / std::mt19937 here are the real ones.
struct Device {
    using result_type = unsigned int;
    std::optional<std::random_device> real;
    Device() { if (romis_shim_asis()) real.emplace(); }
    result_type operator()() { return real ? (*real)() : 0u; }
};
struct Engine {
    using result_type = uint32_t;
    std::optional<std::mt19937> real;
    explicit Engine(result_type seed = 0u) { if (romis_shim_asis()) real.emplace(seed); else romis_shim_engine_ctor(); }
    static constexpr result_type min() { return 0u; }
    static constexpr result_type max() { return 0xffffffffu; }
    result_type operator()() { return real ? (result_type)(*real)() : romis_shim_engine_next(); }
};
// uniform integer in [a, b]: multiply-shift mapping of ONE 32-bit draw (romis_rng.h)
template <class I = int>
struct UniformInt {
    I a, b;
    UniformInt(I a_, I b_) : a(a_), b(b_) {}
    template <class G> I operator()(G& g) {
        uint32_t range = uint32_t(b - a) + 1u;
        return a + I((uint64_t(uint32_t(g())) * uint64_t(range)) >> 32);
    }
};
}
namespace std {
using romis_random_device = ::romis_shim::Device;
using romis_mt19937 = ::romis_shim::Engine;
template <class I = int> using romis_uniform_int_distribution = ::romis_shim::UniformInt<I>;
}
#define random_device romis_random_device
#define mt19937 romis_mt19937
#define uniform_int_distribution romis_uniform_int_distribution
