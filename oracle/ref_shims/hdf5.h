/* TEST INFRASTRUCTURE (oracle/): stand-in for <hdf5.h>.  The reference's MED reader and XDMF writer compile against
 * it; none of their HDF5 paths is on the oracle's path (box / Gambit meshes, no output files): every call aborts. */
#ifndef FEMUS_B200_ORACLE_HDF5_SHIM_H
#define FEMUS_B200_ORACLE_HDF5_SHIM_H
#include <stdio.h>
#include <stdlib.h>
typedef long hid_t;
typedef int herr_t;
typedef int htri_t;
typedef unsigned long long hsize_t;
typedef long long hssize_t;
typedef struct { hsize_t nlinks; } H5G_info_t;
typedef struct { int type; } H5O_info_t;
#define H5P_DEFAULT 0
#define H5S_ALL 0
#define H5T_NATIVE_INT 1
#define H5T_NATIVE_UINT 2
#define H5T_NATIVE_DOUBLE 3
#define H5T_NATIVE_CHAR 4
#define H5F_ACC_RDWR 1
#define H5F_ACC_TRUNC 2
#define H5F_ACC_RDONLY 0
#define H5_ITER_INC 0
#define H5_INDEX_NAME 0
#define H5O_INFO_ALL 0
#define H5I_INVALID_HID (-1)
static inline long femus_b200_no_hdf5_(const char* f) { fprintf(stderr, "oracle hdf5 shim: %s called\n", f); abort(); return -1; }
#define FEMUS_B200_H5STUB(name) template <class... A> static inline long name(A...) { return femus_b200_no_hdf5_(#name); }
FEMUS_B200_H5STUB(H5Fopen) FEMUS_B200_H5STUB(H5Fcreate) FEMUS_B200_H5STUB(H5Fclose)
FEMUS_B200_H5STUB(H5Gopen) FEMUS_B200_H5STUB(H5Gcreate) FEMUS_B200_H5STUB(H5Gclose) FEMUS_B200_H5STUB(H5Gget_info)
FEMUS_B200_H5STUB(H5Dopen) FEMUS_B200_H5STUB(H5Dcreate) FEMUS_B200_H5STUB(H5Dclose) FEMUS_B200_H5STUB(H5Dread) FEMUS_B200_H5STUB(H5Dwrite)
FEMUS_B200_H5STUB(H5Dget_space) FEMUS_B200_H5STUB(H5Screate_simple) FEMUS_B200_H5STUB(H5Sclose) FEMUS_B200_H5STUB(H5Sget_simple_extent_dims)
FEMUS_B200_H5STUB(H5Aopen) FEMUS_B200_H5STUB(H5Aread) FEMUS_B200_H5STUB(H5Aclose)
FEMUS_B200_H5STUB(H5Oget_info) FEMUS_B200_H5STUB(H5Lget_name_by_idx) FEMUS_B200_H5STUB(H5Lexists)
#endif
