/* TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for SUNDIALS'
 * <nvector/nvector_serial.h>.  SUNDIALS is not installed in this image; the
 * reference RHS (src/Model/f.cpp:14-15, src/ModelData/MD_initialize.cpp:117-135,
 * src/ModelData/MD_update.cpp:190-196) needs only NV_DATA_S / NV_Ith_S, so this
 * shim lets the reference's own sources compile unchanged into oracle/_ref/. */
#ifndef SHUD_B200_NVECTOR_SERIAL_SHIM_H
#define SHUD_B200_NVECTOR_SERIAL_SHIM_H
typedef double realtype;
typedef long sunindextype;
struct _N_VectorContent_Serial {
    sunindextype length;
    int own_data;
    realtype *data;
};
typedef struct _N_VectorContent_Serial *N_VectorContent_Serial;
struct _generic_N_Vector {
    void *content;
    void *ops;
    void *sunctx;
};
typedef struct _generic_N_Vector *N_Vector;
#define NV_CONTENT_S(v) ((N_VectorContent_Serial)((v)->content))
#define NV_LENGTH_S(v) (NV_CONTENT_S(v)->length)
#define NV_DATA_S(v) (NV_CONTENT_S(v)->data)
#define NV_Ith_S(v, i) (NV_DATA_S(v)[i])
#endif
