// TEST INFRASTRUCTURE (oracle build only) -- minimal stand-in for "kgd/settings/configfile.h".
// The reference declares its parameter block with Tools' macros (ptreeconfig.h:17-69,
// ptreeconfig.cpp:5-28) and reads values as `Config::name()`, sometimes caching the returned
// reference (phylogenetictree.hpp:257, 817-819): each parameter is therefore a function-local
// static reached through a const-reference accessor, plus a mutable accessor `name_mut()` the
// oracle driver uses to set e.g. rsetSize before constructing a tree.
#ifndef APT_ORACLE_SHIM_CONFIGFILE_H
#define APT_ORACLE_SHIM_CONFIGFILE_H

#include "kgd/utils/utils.h"

namespace config {

enum Verbosity { QUIET, SHOW, PARANOID };

template <typename F>
struct ConfigFile {};

#define DECLARE_PARAMETER(TYPE, NAME)                 \
  static TYPE& NAME##_mut (void);                     \
  static const TYPE& NAME (void) { return NAME##_mut(); }

#define DEFINE_PARAMETER(TYPE, NAME, ...)             \
  TYPE& CFILE::NAME##_mut (void) {                    \
    static TYPE value = TYPE(__VA_ARGS__);            \
    return value;                                     \
  }

#define DECLARE_DEBUG_PARAMETER(TYPE, NAME, VALUE) DECLARE_PARAMETER(TYPE, NAME)
#define DEFINE_DEBUG_PARAMETER(TYPE, NAME, VALUE)  DEFINE_PARAMETER(TYPE, NAME, VALUE)

} // namespace config

#endif
