// STAND-IN for the sol2 (Lua binding) API surface used by PFEM3D's utility/SolTable.hpp.  TEST INFRASTRUCTURE ONLY.
//
// The reference reads its parameters and boundary-condition functions from a Lua file through sol2; neither Lua nor
// sol2 exists in this image.  This header is original code: "tables" are C++ maps of std::any that the test driver
// (oracle/refbuild/ref_driver.cpp) fills with the numbers of the test case and with C++ closures for the BC functions,
// so that the reference's unmodified SolTable.hpp, Equation.cpp and equation constructors compile and run against it.
#pragma once
#include <any>
#include <array>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace sol {

class table;
struct object;
using Args = std::vector<std::any>;
using function = std::function<std::any(const Args&)>;

namespace standin {
template <typename T> struct is_std_array : std::false_type {};
template <typename T, std::size_t N> struct is_std_array<std::array<T, N>> : std::true_type {};
template <typename T> struct is_std_vector : std::false_type {};
template <typename T> struct is_std_vector<std::vector<T>> : std::true_type {};

template <typename T> T convert(const std::any& a) {
    if constexpr (std::is_same<T, bool>::value) {
        if (auto p = std::any_cast<bool>(&a)) return *p;
        throw std::runtime_error("sol stand-in: value is not a bool");
    } else if constexpr (std::is_arithmetic<T>::value) {
        if (auto p = std::any_cast<double>(&a)) return static_cast<T>(*p);
        if (auto p = std::any_cast<int>(&a)) return static_cast<T>(*p);
        if (auto p = std::any_cast<bool>(&a)) return static_cast<T>(*p);
        throw std::runtime_error("sol stand-in: value is not a number");
    } else if constexpr (std::is_same<T, std::string>::value) {
        if (auto p = std::any_cast<std::string>(&a)) return *p;
        throw std::runtime_error("sol stand-in: value is not a string");
    } else if constexpr (is_std_vector<T>::value) {
        if (auto p = std::any_cast<std::vector<double>>(&a)) return T(p->begin(), p->end());
        throw std::runtime_error("sol stand-in: value is not a number list");
    } else if constexpr (is_std_array<T>::value) {
        T out{};
        if (auto p = std::any_cast<std::vector<double>>(&a)) {
            for (std::size_t i = 0; i < out.size() && i < p->size(); ++i) out[i] = (*p)[i];
            return out;
        }
        throw std::runtime_error("sol stand-in: value is not a number list");
    } else {
        if (auto p = std::any_cast<T>(&a)) return *p;
        throw std::runtime_error("sol stand-in: unsupported conversion");
    }
}
}  // namespace standin

struct object {
    std::any v;
    bool valid() const { return v.has_value(); }
    template <typename T> T as() const { return standin::convert<T>(v); }
    template <typename T> T get() const { return standin::convert<T>(v); }
};

struct unsafe_function_result {
    std::any v;
    std::string err;
    bool ok = true;
    bool valid() const { return ok; }
    template <typename T> T get() const { return standin::convert<T>(v); }
};
using protected_function_result = unsafe_function_result;

class error : public std::runtime_error {
  public:
    error(const std::string& w) : std::runtime_error(w) {}
    error(const unsafe_function_result& r) : std::runtime_error(r.err) {}
};

class proxy {
  public:
    proxy(std::any v, std::string key) : v_(std::move(v)), key_(std::move(key)) {}
    bool valid() const { return v_.has_value(); }
    template <typename T> T get() const { return standin::convert<T>(v_); }
    template <typename... A> unsafe_function_result operator()(A&&... a) const {
        unsafe_function_result r;
        const function* f = std::any_cast<function>(&v_);
        if (!f) {
            r.ok = false;
            r.err = "attempt to call a nil value (field '" + key_ + "')";
            return r;
        }
        Args args;
        (args.emplace_back(std::decay_t<A>(std::forward<A>(a))), ...);
        try {
            r.v = (*f)(args);
        } catch (const std::exception& e) {
            r.ok = false;
            r.err = e.what();
        }
        return r;
    }
    const std::any& any() const { return v_; }

  private:
    std::any v_;
    std::string key_;
};

class table {
  public:
    using Fields = std::map<std::string, std::any>;
    table() : f_(std::make_shared<Fields>()) {}
    table(const proxy& p) { *this = p; }
    table(const object& o) { f_ = standin::convert<table>(o.v).f_; }
    table& operator=(const proxy& p) {
        const table* t = std::any_cast<table>(&p.any());
        if (!t) throw std::runtime_error("sol stand-in: value is not a table");
        f_ = t->f_;
        return *this;
    }
    proxy operator[](const std::string& key) const {
        auto it = f_->find(key);
        return proxy(it == f_->end() ? std::any() : it->second, key);
    }
    template <typename V> void set(const std::string& key, V v) { (*f_)[key] = std::any(std::move(v)); }
    void set_function(const std::string& key, function fn) { (*f_)[key] = std::any(std::move(fn)); }
    void for_each(std::function<void(object, object)> fn) const {
        for (auto& kv : *f_) fn(object{std::any(kv.first)}, object{kv.second});
    }

  private:
    std::shared_ptr<Fields> f_;
};

class state : public table {};

}  // namespace sol
