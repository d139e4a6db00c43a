/* Generated-by-hand equivalent of src/version_string.h.in (five @..@ tokens). */
#ifndef SWIFT_VERSION_STRING_H
#define SWIFT_VERSION_STRING_H
#define PACKAGE_VERSION "2026.04"
#define GIT_REVISION "oracle"
#define GIT_BRANCH "oracle"
#define GIT_DATE "unknown"
#define SWIFT_CFLAGS "-O3 -ffp-contract=off"
#endif
