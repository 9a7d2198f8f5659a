// TEST INFRASTRUCTURE ONLY (see ../../cuda_runtime.h).
#pragma once
namespace thrust {
template <class T> struct counting_iterator {
    T base;
    explicit counting_iterator(T b) : base(b) {}
    T operator[](long long i) const { return base + (T)i; }
};
}  // namespace thrust
