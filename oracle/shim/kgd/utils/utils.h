// TEST INFRASTRUCTURE (oracle build only) -- minimal stand-in for kgd/Tools "kgd/utils/utils.h".
// Provides exactly the symbols the reference's tree core uses
// (phylogenetictree.hpp:170,174,951,984,1012; treetypes.h:53-55) and the standard headers the
// reference relies on transitively (std::iota, std::bind, std::function, assert ...).
#ifndef APT_ORACLE_SHIM_UTILS_H
#define APT_ORACLE_SHIM_UTILS_H

#include <sys/types.h>   // uint

#include <algorithm>
#include <cassert>
#include <cmath>
#include <filesystem>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <numeric>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

namespace stdfs = std::filesystem;

namespace utils {

/// Concatenates its arguments into a message and throws it as an exception of type E
template <typename E = std::invalid_argument, typename... ARGS>
[[noreturn]] void Thrower (ARGS... args) {
  std::ostringstream oss;
  (oss << ... << args);
  throw E(oss.str());
}

template <typename T>
std::string className (void) { return typeid(T).name(); }

inline std::string readAll (const std::string &path) {
  std::ifstream ifs (path);
  std::stringstream ss;
  ss << ifs.rdbuf();
  return ss.str();
}

/// Unnormalised gaussian, the compatibility kernel documented at crossover.h:76-86
inline double gauss (double x, double mu, double sigma) {
  return std::exp(-((x - mu) * (x - mu)) / (2. * sigma * sigma));
}

} // namespace utils

#endif
