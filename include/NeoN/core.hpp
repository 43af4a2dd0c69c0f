// NeoN core for the B200 build: primitives, the (single) GPU executor, device Vector/View, Dictionary/TokenList.
//
// Same names and meaning as the reference (src/NeoN/include/NeoN/core/{primitives,executor,vector,view,dictionary,
// tokenList,input,error}.hpp), but there is exactly one executor -- a CUDA device + stream -- so there is no
// std::variant dispatch, no Kokkos, and no CPU fallback: every operation goes to libfvk (include/fvk.h).
#pragma once

#include "fvk.h"

#include <any>
#include <array>
#include <cmath>
#include <cstdint>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace NeoN
{
// ---- primitives (core/primitives/{scalar,label,vec3}.hpp) -------------------------------------------
using scalar = double;
using label = int32_t;
using localIdx = int32_t;
using globalIdx = int64_t;
constexpr scalar ROOTVSMALL = 1e-18;

class Vec3
{
public:

    Vec3() : cmpts_ {0.0, 0.0, 0.0} {}
    Vec3(scalar x, scalar y, scalar z) : cmpts_ {x, y, z} {}
    explicit Vec3(scalar v) : cmpts_ {v, v, v} {}
    scalar* data() { return cmpts_; }
    const scalar* data() const { return cmpts_; }
    constexpr size_t size() const { return 3; }
    scalar& operator[](size_t i) { return cmpts_[i]; }
    scalar operator[](size_t i) const { return cmpts_[i]; }
    scalar& operator()(size_t i) { return cmpts_[i]; }
    scalar operator()(size_t i) const { return cmpts_[i]; }
    bool operator==(const Vec3& r) const { return cmpts_[0] == r(0) && cmpts_[1] == r(1) && cmpts_[2] == r(2); }
    Vec3 operator+(const Vec3& r) const { return Vec3(cmpts_[0] + r(0), cmpts_[1] + r(1), cmpts_[2] + r(2)); }
    Vec3 operator-(const Vec3& r) const { return Vec3(cmpts_[0] - r(0), cmpts_[1] - r(1), cmpts_[2] - r(2)); }
    Vec3 operator*(const scalar& r) const { return Vec3(cmpts_[0] * r, cmpts_[1] * r, cmpts_[2] * r); }
    Vec3& operator+=(const Vec3& r) { return *this = *this + r; }
    Vec3& operator-=(const Vec3& r) { return *this = *this - r; }
    Vec3& operator*=(const scalar& r) { return *this = *this * r; }

private:

    scalar cmpts_[3];
};
inline Vec3 operator*(const scalar& s, Vec3 r) { r *= s; return r; }
inline scalar operator&(const Vec3& l, Vec3 r) { return l[0] * r[0] + l[1] * r[1] + l[2] * r[2]; }
inline scalar mag(const Vec3& v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
inline std::ostream& operator<<(std::ostream& o, const Vec3& v) { return o << "(" << v[0] << " " << v[1] << " " << v[2] << ")"; }
static_assert(sizeof(Vec3) == 24, "Vec3 is 3 contiguous doubles (AoS), the layout contract of the C ABI");

template<typename T> inline T one();
template<> inline scalar one<scalar>() { return 1.0; }
template<> inline Vec3 one<Vec3>() { return Vec3(1.0, 1.0, 1.0); }
template<typename T> inline T zero();
template<> inline scalar zero<scalar>() { return 0.0; }
template<> inline localIdx zero<localIdx>() { return 0; }
template<> inline Vec3 zero<Vec3>() { return Vec3(0.0, 0.0, 0.0); }
template<typename T> constexpr int nComponents() { return std::is_same_v<T, Vec3> ? 3 : 1; }

// ---- errors (core/error.hpp): NF_ERROR_EXIT / NF_THROW keep their names; both throw ---------------------
class NeoNException : public std::runtime_error
{
public:
    explicit NeoNException(const std::string& m) : std::runtime_error(m) {}
};
#define NF_THROW(msg) throw ::NeoN::NeoNException(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + std::string(msg))
#define NF_ERROR_EXIT(msg) NF_THROW(msg)
#define NF_ASSERT(cond, msg) do { if (!(cond)) NF_THROW(msg); } while (0)
inline void check(int rc)
{
    if (rc != FVK_OK) throw NeoNException(std::string("libfvk: ") + fvk_last_error());
}

// ---- executor (core/executor/{executor,GPUExecutor}.hpp) ----------------------------------------------
// The reference's Executor is std::variant<SerialExecutor, CPUExecutor, GPUExecutor>; here GPUExecutor is the
// only alternative: a CUDA device and the stream every kernel of this executor is launched on.
class GPUExecutor
{
public:
    explicit GPUExecutor(int device = 0, fvk_stream stream = nullptr) : device_(device), stream_(stream)
    {
        int n = 0;
        check(fvk_device_count(&n));
        if (device >= n) NF_THROW("GPUExecutor: no such CUDA device (there is no CPU fallback)");
        check(fvk_set_device(device));
    }
    std::string name() const { return "GPUExecutor"; }
    int device() const { return device_; }
    fvk_stream stream() const { return stream_; }
    void sync() const { check(fvk_stream_sync(stream_)); }
    bool operator==(const GPUExecutor& r) const { return device_ == r.device_ && stream_ == r.stream_; }
private:
    int device_;
    fvk_stream stream_;
};
using Executor = GPUExecutor;

// ---- View / Vector (core/view.hpp, core/vector/vector.hpp) ---------------------------------------------------
template<typename T>
struct View
{
    T* ptr = nullptr;
    size_t n = 0;
    T* data() const { return ptr; }
    size_t size() const { return n; }
};

template<typename T>
class Vector
{
public:
    using ElementType = T;
    Vector(const Executor& exec, size_t n) : exec_(exec), n_(n) { alloc(); }
    Vector(const Executor& exec, size_t n, T value) : exec_(exec), n_(n) { alloc(); fillWith(value); }
    Vector(const Executor& exec, const std::vector<T>& host) : exec_(exec), n_(host.size()) { alloc(); copyFromHost(host.data()); }
    Vector(const Vector& r) : exec_(r.exec_), n_(r.n_)
    {
        alloc();
        check(fvk_memcpy_d2d(data_, r.data_, n_ * sizeof(T), exec_.stream()));
    }
    Vector(Vector&& r) noexcept : exec_(r.exec_), n_(r.n_), data_(r.data_) { r.data_ = nullptr; r.n_ = 0; }
    Vector& operator=(const Vector& r)
    { // setContainer (core/containerFreeFunctions.hpp:101-115)
        if (this == &r) return *this;
        if (n_ != r.n_) { release(); n_ = r.n_; alloc(); }
        check(fvk_memcpy_d2d(data_, r.data_, n_ * sizeof(T), exec_.stream()));
        return *this;
    }
    ~Vector() { release(); }

    const Executor& exec() const { return exec_; }
    size_t size() const { return n_; }
    localIdx ssize() const { return static_cast<localIdx>(n_); }
    T* data() { return data_; }
    const T* data() const { return data_; }
    View<T> view() { return {data_, n_}; }
    View<const T> view() const { return {data_, n_}; }
    double* raw() { return reinterpret_cast<double*>(data_); }
    const double* raw() const { return reinterpret_cast<const double*>(data_); }

    std::vector<T> copyToHost() const
    {
        std::vector<T> h(n_);
        if (n_) { check(fvk_memcpy_d2h(h.data(), data_, n_ * sizeof(T), exec_.stream())); exec_.sync(); }
        return h;
    }
    void copyFromHost(const T* h)
    {
        if (n_) { check(fvk_memcpy_h2d(data_, h, n_ * sizeof(T), exec_.stream())); exec_.sync(); }
    }
    // vectorFreeFunctions.cpp:18-106
    Vector& operator+=(const Vector& r) { sameSize(r); if constexpr (isFp()) check(fvk_vec_add(nd(), raw(), r.raw(), exec_.stream())); return *this; }
    Vector& operator-=(const Vector& r) { sameSize(r); if constexpr (isFp()) check(fvk_vec_sub(nd(), raw(), r.raw(), exec_.stream())); return *this; }
    Vector& operator*=(scalar a) { if constexpr (isFp()) check(fvk_vec_scale(nd(), a, raw(), exec_.stream())); return *this; }

    void fillWith(T value)
    {
        if (!n_) return;
        if constexpr (std::is_same_v<T, scalar>) check(fvk_vec_fill(nd(), value, raw(), exec_.stream()));
        else if constexpr (std::is_same_v<T, Vec3>)
        {
            if (value[0] == value[1] && value[1] == value[2]) check(fvk_vec_fill(nd(), value[0], raw(), exec_.stream()));
            else { std::vector<T> h(n_, value); copyFromHost(h.data()); }
        }
        else { std::vector<T> h(n_, value); copyFromHost(h.data()); }
    }

private:
    static constexpr bool isFp() { return std::is_same_v<T, scalar> || std::is_same_v<T, Vec3>; }
    int64_t nd() const { return int64_t(n_) * nComponents<T>(); }
    void sameSize(const Vector& r) const { if (n_ != r.n_) NF_THROW("Vector size mismatch"); }
    void alloc()
    {
        data_ = nullptr;
        if (n_) { void* p = nullptr; check(fvk_malloc(&p, n_ * sizeof(T))); data_ = static_cast<T*>(p); }
    }
    void release() { if (data_) fvk_free(data_); data_ = nullptr; }
    Executor exec_;
    size_t n_ = 0;
    T* data_ = nullptr;
};

template<typename T> inline void fill(Vector<T>& v, T value) { v.fillWith(value); } // containerFreeFunctions.hpp:51-64
template<typename T> inline void setContainer(Vector<T>& dst, const Vector<T>& src) { dst = src; }
inline void scalarMul(Vector<scalar>& v, scalar a) { v *= a; }
inline void add(Vector<scalar>& a, const Vector<scalar>& b) { a += b; }
inline void sub(Vector<scalar>& a, const Vector<scalar>& b) { a -= b; }


// ---- RuntimeSelectionFactory (core/runtimeSelectionFactory.hpp:193-413): plugins register by NAME --------------------
// `class MyScheme : public Base::template Register<MyScheme>` gives the class the reference's shape (static name(), doc(),
// schema()); NF_REGISTER(Base, MyScheme) -- or Base::template Register<MyScheme>::add() -- enters it into Base's table under
// MyScheme::name(). The reference does the same with a static initialiser inside Register (:406-412), which needs explicit
// template instantiations in its .cpp files; this build is header-only, so the registration is an inline variable.
template<typename... Args>
struct Parameters
{
};
template<typename Base, typename Params>
class RuntimeSelectionFactory;
template<typename Base, typename... Args>
class RuntimeSelectionFactory<Base, Parameters<Args...>>
{
public:
    using CreatorFunc = std::function<std::unique_ptr<Base>(Args...)>;
    using LookupTable = std::map<std::string, CreatorFunc>;
    static LookupTable& table()
    {
        static LookupTable tbl;
        return tbl;
    }
    static std::vector<std::string> entries()
    {
        std::vector<std::string> e;
        for (const auto& kv : table()) e.push_back(kv.first);
        return e;
    }
    static size_t size() { return table().size(); }
    static bool contains(const std::string& key) { return table().count(key) > 0; }
    static void keyExistsOrError(const std::string& key)
    {
        if (contains(key)) return;
        std::string known;
        for (const auto& kv : table()) known += " " + kv.first;
        NF_ERROR_EXIT("Could not find constructor for " + key + ". Valid constructors are:" + known);
    }
    static std::unique_ptr<Base> create(const std::string& key, Args... args)
    {
        keyExistsOrError(key);
        return table().at(key)(std::forward<Args>(args)...);
    }
    template<typename Derived>
    class Register : public Base
    {
    public:
        using Base::Base;
        static bool add()
        {
            RuntimeSelectionFactory::table()[Derived::name()] = [](Args... a) { return std::unique_ptr<Base>(new Derived(std::forward<Args>(a)...)); };
            return true;
        }
    };
    virtual ~RuntimeSelectionFactory() = default;
};
#define NF_REGISTER_CAT2(a, b) a##b
#define NF_REGISTER_CAT(a, b) NF_REGISTER_CAT2(a, b)
// namespace-scope registration: NF_REGISTER((fvcc::SurfaceInterpolationFactory<scalar>), (MyScheme<scalar>))
#define NF_UNPAREN(...) __VA_ARGS__
#define NF_REGISTER(Base, Derived) \
    inline const bool NF_REGISTER_CAT(nf_registered_, __COUNTER__) = NF_UNPAREN Base ::template Register<NF_UNPAREN Derived>::add()

// ---- Dictionary / TokenList / Input (core/{dictionary,tokenList,input}.hpp) -- host-only configuration ------------
class Dictionary
{
public:
    Dictionary() = default;
    Dictionary(std::initializer_list<std::pair<const std::string, std::any>> init) : data_(init) {}
    void insert(const std::string& key, const std::any& value) { data_[key] = value; }
    bool contains(const std::string& key) const { return data_.count(key) > 0; }
    void remove(const std::string& key) { data_.erase(key); }
    bool isDict(const std::string& key) const
    {
        auto it = data_.find(key);
        return it != data_.end() && it->second.type() == typeid(Dictionary);
    }
    template<typename T> T& get(const std::string& key)
    {
        auto it = data_.find(key);
        if (it == data_.end()) NF_THROW("Key " + key + " not found in dictionary");
        try { return std::any_cast<T&>(it->second); }
        catch (const std::bad_any_cast&) { NF_THROW("Bad type for key " + key); }
    }
    template<typename T> const T& get(const std::string& key) const { return const_cast<Dictionary*>(this)->get<T>(key); }
    template<typename T> T getOr(const std::string& key, T dflt) const { return contains(key) ? get<T>(key) : dflt; }
    Dictionary& subDict(const std::string& key) { return get<Dictionary>(key); }
    const Dictionary& subDict(const std::string& key) const { return get<Dictionary>(key); }
    std::any& operator[](const std::string& key) { return data_[key]; }
    std::vector<std::string> keys() const { std::vector<std::string> k; for (auto& e : data_) k.push_back(e.first); return k; }
private:
    std::map<std::string, std::any> data_;
};

class TokenList
{
public:
    TokenList() = default;
    TokenList(std::initializer_list<std::string> t) : data_(t) {}
    explicit TokenList(std::vector<std::string> t) : data_(std::move(t)) {}
    size_t size() const { return data_.size(); }
    bool empty() const { return data_.empty(); }
    const std::string& operator[](size_t i) const { if (i >= data_.size()) NF_THROW("TokenList index out of range"); return data_[i]; }
    std::string popFront() { if (data_.empty()) NF_THROW("TokenList is empty"); std::string s = data_.front(); data_.erase(data_.begin()); return s; }
    static TokenList split(const std::string& s)
    {
        std::vector<std::string> t; std::string cur;
        for (char ch : s) { if (ch == ' ' || ch == '\t') { if (!cur.empty()) t.push_back(cur); cur.clear(); } else cur.push_back(ch); }
        if (!cur.empty()) t.push_back(cur);
        return TokenList(t);
    }
private:
    std::vector<std::string> data_;
};
using Input = TokenList;

} // namespace NeoN
