/* oracle/shims/suitesparse/umfpack.h — TEST INFRASTRUCTURE ONLY.
 *
 * UMFPACK (SuiteSparse, version not pinned by the reference: Makefile:5, src/fields.cpp:2,
 * src/fields3d.cpp:1) is not in /root/reference and not installed here.  The reference only uses it
 * as "factor once, solve A^T x = b each step" (src/fields.cpp:273-275,311,347,352,364), and the
 * matrix is fully specified by src/fields.cpp:133-259, so any accurate direct solve is a valid
 * stand-in.  umfpack_shim.cpp implements these entry points with a banded LU (no pivoting is
 * needed: every row is either an identity row or a weakly diagonally dominant stencil row).
 * Parity at this third-party boundary is therefore "unpinned" in the task's sense; the operator
 * itself (what is being solved) is pinned against the reference's own matrix build.
 */
#ifndef MAG2D_ORACLE_UMFPACK_SHIM_H
#define MAG2D_ORACLE_UMFPACK_SHIM_H
#ifdef __cplusplus
extern "C" {
#endif
#define UMFPACK_A 0
#define UMFPACK_At 1
#define UMFPACK_OK 0
#define UMFPACK_ERROR_out_of_memory (-1)
#define UMFPACK_ERROR_invalid_Numeric_object (-3)
#define UMFPACK_ERROR_file_IO (-17)
typedef long UF_long;
int umfpack_di_symbolic(int n_row, int n_col, const int Ap[], const int Ai[], const double Ax[],
                        void** Symbolic, const double Control[], double Info[]);
int umfpack_di_numeric(const int Ap[], const int Ai[], const double Ax[], void* Symbolic,
                       void** Numeric, const double Control[], double Info[]);
int umfpack_di_solve(int sys, const int Ap[], const int Ai[], const double Ax[], double X[],
                     const double B[], void* Numeric, const double Control[], double Info[]);
void umfpack_di_free_symbolic(void** Symbolic);
void umfpack_di_free_numeric(void** Numeric);
int umfpack_di_save_numeric(void* Numeric, char* filename);
int umfpack_di_load_numeric(void** Numeric, char* filename);
#ifdef __cplusplus
}
#endif
#endif
