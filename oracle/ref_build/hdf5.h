/* Stub <hdf5.h>: HDF5 is not installed; the reference's
 * src/lightcone/lightcone_particle_io.h:27 includes it unconditionally.
 * Only the typedef names are needed to parse the headers. */
#ifndef ORACLE_STUB_HDF5_H
#define ORACLE_STUB_HDF5_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
typedef long hid_t;
typedef unsigned long long hsize_t;
typedef int herr_t;
typedef int htri_t;
#endif
