/* Force-included (-include) when compiling the unmodified reference sources without Boost:
 * message.h:52 names boost::iostreams::filtering_streambuf outside its #if, and
 * covariancematrix.cpp:1315 uses unqualified isinf/isnan. */
#ifndef DGRM_ORACLE_SHIM_FWD_H
#define DGRM_ORACLE_SHIM_FWD_H
#ifdef __cplusplus
#include <cmath>
#include <limits>   /* communicator.cpp:582 uses std::numeric_limits without including <limits> */
#include <cstdlib>
#include <cstring>
#include <algorithm>
namespace boost { namespace iostreams {
template <typename Mode, typename Ch = char> class filtering_streambuf;
struct input; struct output;
} }
using std::isinf;
using std::isnan;
#endif
#endif
