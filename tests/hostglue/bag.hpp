// Named-array bag used to move flat arrays between the C++ host-glue stages and Python (ctypes/numpy).
#pragma once
#include <array>
#include <complex>
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace hg {

enum DType : int { F64 = 0, I32 = 1, C128 = 2, I64 = 3 };

struct Arr {
    int dtype = F64;
    int64_t n = 0;
    std::vector<char> data;
};

inline size_t dtype_size(int dt) { return dt == F64 ? 8 : dt == I32 ? 4 : dt == C128 ? 16 : 8; }

struct Bag {
    std::map<std::string, Arr> arrs;

    void put_raw(const std::string& name, int dtype, const void* p, int64_t n) {
        Arr& a = arrs[name];
        a.dtype = dtype;
        a.n = n;
        a.data.resize((size_t)n * dtype_size(dtype));
        if (n) std::memcpy(a.data.data(), p, a.data.size());
    }
    void put(const std::string& name, const std::vector<double>& v) { put_raw(name, F64, v.data(), (int64_t)v.size()); }
    void put(const std::string& name, const std::vector<int>& v) { put_raw(name, I32, v.data(), (int64_t)v.size()); }
    void put(const std::string& name, const std::vector<std::complex<double>>& v) { put_raw(name, C128, v.data(), (int64_t)v.size()); }
    void put_scalar(const std::string& name, double v) { put_raw(name, F64, &v, 1); }
    void put_scalar_i(const std::string& name, int v) { put_raw(name, I32, &v, 1); }
    template <class T, size_t K>
    void put(const std::string& name, const std::vector<std::array<T, K>>& v) {
        put_raw(name, std::is_same<T, int>::value ? I32 : F64, v.data(), (int64_t)(v.size() * K));
    }

    const Arr& get(const std::string& name) const {
        auto it = arrs.find(name);
        if (it == arrs.end()) throw std::runtime_error("bag: missing array '" + name + "'");
        return it->second;
    }
    template <class T>
    std::vector<T> get_vec(const std::string& name, int dtype) const {
        const Arr& a = get(name);
        if (a.dtype != dtype) throw std::runtime_error("bag: dtype mismatch for '" + name + "'");
        std::vector<T> v((size_t)a.n);
        if (a.n) std::memcpy(v.data(), a.data.data(), a.data.size());
        return v;
    }
    std::vector<double> get_f64(const std::string& n) const { return get_vec<double>(n, F64); }
    std::vector<int> get_i32(const std::string& n) const { return get_vec<int>(n, I32); }
    std::vector<std::complex<double>> get_c128(const std::string& n) const { return get_vec<std::complex<double>>(n, C128); }
};

}  // namespace hg
