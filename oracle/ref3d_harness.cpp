// oracle/ref3d_harness.cpp — TEST INFRASTRUCTURE ONLY.
// extern "C" access to the parts of the reference's 3-D code that still compile: Field3D (deposit, interpolate,
// grad; src/Field3D.hpp), Geometry and Solver (src/fields3d.hpp, src/fields3d.cpp) with the UMFPACK shim.  The
// sources are compiled where they lie under /root/reference (oracle/Makefile); nothing is copied.
// Field3D leaves xmax/ymax/zmax uninitialised; the harness sets them to (imax-1)/idx etc.
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include <GetPot>
#define private public
#define protected public
#include "fields3d.hpp"
#undef private
#undef protected

struct Ref3
{
    GetPot* config;
    Param* param;
    Geometry* geometry;
    Solver* solver;
    Field3D* u;
    Field3D* rho;
    Field3D* voltage;     // ElMag3D::voltage: reset, then every electrode's set_voltage (fields3d.cpp:176-179)
    std::string error;
};

static void fix_extent(Field3D& f)
{
    f.xmax = (f.imax - 1) / f.idx;
    f.ymax = (f.jmax - 1) / f.idy;
    f.zmax = (f.kmax - 1) / f.idz;
}

extern "C" {

void* ref3_create(const char* config_file, int with_solver)
{
    Ref3* h = new Ref3();
    try
    {
        h->config = new GetPot(config_file);
        h->param = new Param(*h->config);
        Param& p = *h->param;
        h->geometry = new Geometry(p);
        h->solver = with_solver ? new Solver(*h->geometry, p) : nullptr;
        h->u = new Field3D(p.x_sampl, p.y_sampl, p.z_sampl, p.dx, p.dy, p.dz);
        h->rho = new Field3D(p.x_sampl, p.y_sampl, p.z_sampl, p.dx, p.dy, p.dz);
        h->voltage = new Field3D(p.x_sampl, p.y_sampl, p.z_sampl, p.dx, p.dy, p.dz);
        fix_extent(*h->u);
        fix_extent(*h->rho);
        h->u->reset();
        h->rho->reset();
        h->voltage->reset();
        for (unsigned int i = 0; i < h->geometry->electrodes.size(); i++)
            h->geometry->electrodes[i]->set_voltage(*h->voltage, h->geometry->mask);
    }
    catch (std::exception& e)
    {
        h->error = e.what();
    }
    return h;
}
const char* ref3_error(void* hv) { return ((Ref3*)hv)->error.c_str(); }
void ref3_destroy(void* hv) { delete (Ref3*)hv; }
void ref3_dims(void* hv, int* dims, double* d)
{
    Param& p = *((Ref3*)hv)->param;
    dims[0] = p.x_sampl; dims[1] = p.y_sampl; dims[2] = p.z_sampl;
    d[0] = p.dx; d[1] = p.dy; d[2] = p.dz; d[3] = p.idx; d[4] = p.idy; d[5] = p.idz;
    d[6] = p.x_max; d[7] = p.y_max; d[8] = p.z_max; d[9] = p.macroparticle_factor;
}
void ref3_mask(void* hv, signed char* mask, double* voltage)
{
    Ref3* h = (Ref3*)hv;
    const size_t n = (size_t)h->geometry->mask.imax * h->geometry->mask.jmax * h->geometry->mask.kmax;
    memcpy(mask, h->geometry->mask[0][0], n);
    memcpy(voltage, (*h->voltage)[0][0], n * sizeof(double));
}
int ref3_is_free(void* hv, int n, const double* x, const double* y, const double* z, int* out)
{
    Ref3* h = (Ref3*)hv;
    for (int p = 0; p < n; p++) out[p] = h->geometry->is_free(x[p], y[p], z[p]) ? 1 : 0;
    return 0;
}
int ref3_accumulate(void* hv, int n, double charge, const double* x, const double* y, const double* z)
{
    Ref3* h = (Ref3*)hv;
    int thrown = 0;
    for (int p = 0; p < n; p++)
        try { h->rho->accumulate(charge, x[p], y[p], z[p]); }
        catch (std::exception&) { thrown++; }
    return thrown;
}
void ref3_get(void* hv, int which, double* out)
{
    Ref3* h = (Ref3*)hv;
    Field3D& f = which ? *h->rho : *h->u;
    memcpy(out, f[0][0], sizeof(double) * (size_t)f.imax * f.jmax * f.kmax);
}
void ref3_set(void* hv, int which, const double* in)
{
    Ref3* h = (Ref3*)hv;
    Field3D& f = which ? *h->rho : *h->u;
    memcpy(f[0][0], in, sizeof(double) * (size_t)f.imax * f.jmax * f.kmax);
}
int ref3_grad(void* hv, int n, const double* x, const double* y, const double* z, double* gx, double* gy, double* gz, double* val)
{
    Ref3* h = (Ref3*)hv;
    for (int p = 0; p < n; p++)
    {
        h->u->grad(x[p], y[p], z[p], gx[p], gy[p], gz[p]);
        try { val[p] = h->u->interpolate(x[p], y[p], z[p]); }
        catch (std::exception&) { val[p] = NAN; }
    }
    return 0;
}
int ref3_interpolate(void* hv, int n, const double* x, const double* y, const double* z, double* val)
{
    Ref3* h = (Ref3*)hv;
    for (int p = 0; p < n; p++)
        try { val[p] = h->u->interpolate(x[p], y[p], z[p]); }
        catch (std::exception&) { val[p] = NAN; }
    return 0;
}
// Solver::solve(u, voltage, rho): rho is scaled in place into the right-hand side (fields3d.cpp:83-95)
int ref3_solve(void* hv)
{
    Ref3* h = (Ref3*)hv;
    if (!h->solver) return 1;
    h->solver->solve(*h->u, *h->voltage, *h->rho);
    return 0;
}
}
