// TEST INFRASTRUCTURE ONLY (see ../../cuda_runtime.h).
#pragma once
namespace thrust {
template <class Fn, class It, class T = void> struct transform_iterator {
    It it;
    Fn fn;
    transform_iterator(It i, Fn f) : it(i), fn(f) {}
    auto operator[](long long i) const { return fn(it[i]); }
};
}  // namespace thrust
