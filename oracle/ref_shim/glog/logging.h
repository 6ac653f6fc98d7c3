// ref_shim/glog/logging.h — TEST INFRASTRUCTURE (oracle/_ref build only): LOG(x) << ... swallows its arguments.
#pragma once
#include <ostream>
namespace nav24_ref_shim { struct NullLog { template <class T> NullLog& operator<<(const T&) { return *this; } NullLog& operator<<(std::ostream& (*)(std::ostream&)) { return *this; } }; }
#define LOG(sev) ::nav24_ref_shim::NullLog()
#define DLOG(sev) ::nav24_ref_shim::NullLog()
#define VLOG(n) ::nav24_ref_shim::NullLog()
#define DVLOG(n) ::nav24_ref_shim::NullLog()
