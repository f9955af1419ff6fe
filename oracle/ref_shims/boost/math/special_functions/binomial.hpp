// TEST INFRASTRUCTURE: stand-in for <boost/math/special_functions/binomial.hpp> (src/uq/uq.cpp:639)
#pragma once
namespace boost { namespace math {
template <class T> inline T binomial_coefficient(unsigned n, unsigned k) {
  T r = 1;
  for (unsigned i = 1; i <= k; i++) r = r * (T)(n - k + i) / (T)i;
  return r;
}
} }
