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

namespace romis_shim {
struct Device {
    using result_type = unsigned int;
    result_type operator()() { return 0u; }
};
struct Engine {
    using result_type = uint32_t;
    explicit Engine(result_type = 0u) { romis_shim_engine_ctor(); }
    static constexpr result_type min() { return 0u; }
    static constexpr result_type max() { return 0xffffffffu; }
    result_type operator()() { return romis_shim_engine_next(); }
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
