// cli_common.h -- argument-value parsing shared by the three drop-in binaries.  TCLAP (the reference's parser) rejects
// anything that is not entirely a number of the argument's type with "Couldn't read argument value from string"
// (SibeliaZ-LCB/sibeliaz.cpp:37-111, TwoPaCo/src/graphconstructor/constructor.cpp:58-143): so do these.
#pragma once
#include <cerrno>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace cli {

// digits only (no sign, no blanks), no overflow of the destination type
inline bool ParseU64(const char *s, uint64_t &out)
{
    if (!s || *s < '0' || *s > '9') return false;
    char *end = nullptr;
    errno = 0;
    const unsigned long long v = strtoull(s, &end, 10);
    if (*end || errno == ERANGE) return false;
    out = (uint64_t)v;
    return true;
}

inline bool ParseUnsigned(const char *s, unsigned &out)
{
    uint64_t v = 0;
    if (!ParseU64(s, v) || v > UINT_MAX) return false;
    out = (unsigned)v;
    return true;
}

inline bool ParseInt(const char *s, int &out) // optional leading '-'
{
    const bool neg = s && *s == '-';
    uint64_t v = 0;
    if (!ParseU64(neg ? s + 1 : s, v) || v > (uint64_t)INT_MAX) return false;
    out = neg ? -(int)v : (int)v;
    return true;
}

inline void BadValue(const char *value, const char *arg)
{
    fprintf(stderr, "error: Couldn't read argument value from string '%s' for arg %s\n", value ? value : "", arg);
}

} // namespace cli
