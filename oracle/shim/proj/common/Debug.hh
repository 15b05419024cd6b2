// Shadows common/Debug.hh of the reference (which needs Boost.Thread/DateTime): same macro names, assertions
// print to stderr and abort (the reference's default build keeps them live), dev traces compile to nothing.
#ifndef iSAAC_LOG_THREAD_TIMESTAMP_HH
#define iSAAC_LOG_THREAD_TIMESTAMP_HH
#include <iostream>
#include <cstdlib>
// the reference's progress / warning messages go nowhere: a stream without a buffer drops everything before any formatting
namespace isaac_shim { inline std::ostream &quietLog() { static std::ostream quiet(nullptr); return quiet; } }
#define ISAAC_THREAD_CERR isaac_shim::quietLog()
#define ISAAC_ASSERT_MSG(expr, msg) {if (expr) {} else \
{ std::cerr << "ERROR: ***** Internal Program Error - assertion (" << #expr << ") failed in " \
    << __FILE__ << '(' << __LINE__ << "): " << msg << std::endl; ::abort();}}
#define ISAAC_THREAD_CERR_DEV_TRACE(blah)
#define ISAAC_THREAD_CERR_DEV_TRACE_CLUSTER_ID(clusterId, blah)
#define ISAAC_DEV_TRACE_BLOCK(block)
#define ISAAC_TRACE_STAT(prefix)
#endif
