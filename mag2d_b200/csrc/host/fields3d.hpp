// fields3d.hpp — host mirrors of the reference's 3-D field classes on top of libmag2d_b200:
//   Field3D   (src/Field3D.hpp, src/Field3D.cpp: contiguous [i][j][k] block, print / print_vtk formats)
//   Geometry  (src/fields3d.hpp:38-58, src/fields3d.cpp:13-37: box frame + the one-node Quadrupole electrode)
//   ElMag3D   (src/fields3d.hpp:83-121: u, rho, voltage; solve() forwards to mag2d_solve)
#pragma once
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/mag2d_b200.h"
#include "param.hpp"

class Field3D
{
  public:
    int imax, jmax, kmax;
    Field3D(int x_sampl, int y_sampl, int z_sampl, double dx, double dy, double dz, double xmin_ = 0, double ymin_ = 0, double zmin_ = 0)
        : imax(x_sampl), jmax(y_sampl), kmax(z_sampl), base((size_t)x_sampl * y_sampl * z_sampl, 0.0), xmin(xmin_), ymin(ymin_), zmin(zmin_),
          idx(1.0 / dx), idy(1.0 / dy), idz(1.0 / dz)
    {
    }
    double& operator()(int i, int j, int k) { return base[((size_t)i * jmax + j) * kmax + k]; }
    double operator()(int i, int j, int k) const { return base[((size_t)i * jmax + j) * kmax + k]; }
    double* data() { return base.data(); }
    size_t size() const { return base.size(); }
    void reset() { std::fill(base.begin(), base.end(), 0.0); }
    // Field3D::print (Field3D.cpp:14-27): "x y z value" rows, blank line after every k row and every j plane
    void print(std::ostream& out, double factor = 1.0) const
    {
        const double dx = 1.0 / idx, dy = 1.0 / idy, dz = 1.0 / idz;
        for (int i = 0; i < imax; i++)
        {
            for (int j = 0; j < jmax; j++)
            {
                for (int k = 0; k < kmax; k++) out << i * dx << "\t" << j * dy << "\t" << k * dz << "\t" << (*this)(i, j, k) * factor << std::endl;
                out << std::endl;
            }
            out << std::endl;
        }
    }
    // Field3D::print_vtk (Field3D.cpp:37-62): legacy VTK structured points
    void print_vtk(std::ostream& out) const
    {
        const double dx = 1.0 / idx, dy = 1.0 / idy, dz = 1.0 / idz;
        out << "# vtk DataFile Version 2.0\n";
        out << "3D scalar field saved by plasma2d\n";
        out << "ASCII\n";
        out << "DATASET STRUCTURED_POINTS\n";
        out << "DIMENSIONS " << imax << " " << jmax << " " << kmax << std::endl;
        out << "ORIGIN " << xmin << " " << ymin << " " << zmin << std::endl;
        out << "SPACING " << dx << " " << dy << " " << dz << std::endl;
        out << "POINT_DATA " << (long)imax * jmax * kmax << std::endl;
        out << "SCALARS ScalarField double 1\n";
        out << "LOOKUP_TABLE default\n";
        for (size_t m = 0; m < base.size(); m++) out << base[m] << std::endl;
    }
    void print(const char* filename, const std::string& format = "table", double factor = 1.0) const
    {
        std::ofstream out(filename);
        if (format == "table") print(out, factor);
        else if (format == "vtk") print_vtk(out);
        else throw std::runtime_error("Field3D::print() unknown format " + format + "\n");
    }

  private:
    std::vector<double> base;
    double xmin, ymin, zmin, idx, idy, idz;
};

class Geometry
{
  public:
    int x_sampl, y_sampl, z_sampl;
    std::vector<signed char> mask;       // FIXED 0, FREE 2, electrode ids < 0
    explicit Geometry(const Param& p) : x_sampl(p.x_sampl), y_sampl(p.y_sampl), z_sampl(p.z_sampl), mask((size_t)p.x_sampl * p.y_sampl * p.z_sampl)
    {
        for (int i = 0; i < x_sampl; i++)
            for (int j = 0; j < y_sampl; j++)
                for (int k = 0; k < z_sampl; k++)
                    // the reference tests k == z_sampl (fields3d.cpp:28), which never fires: the last z face stays FREE
                    at(i, j, k) = (i == 0 || i == x_sampl - 1 || j == 0 || j == y_sampl - 1 || k == 0 || k == z_sampl) ? MAG2D_FIXED : MAG2D_FREE;
        at(x_sampl / 2, y_sampl / 2, z_sampl / 2) = -1;      // Quadrupole(-1)::set_mask
    }
    signed char& at(int i, int j, int k) { return mask[((size_t)i * y_sampl + j) * z_sampl + k]; }
    // Electrode::set_voltage for every electrode (one, id -1, 1 V)
    void set_voltage(Field3D& voltage) const
    {
        for (size_t m = 0; m < mask.size(); m++)
            if (mask[m] == -1) voltage.data()[m] = 1.0;
    }
};

class ElMag3D
{
  public:
    Field3D u, rho, voltage;
    Geometry geometry;
    mag2d_ctx* gpu = nullptr;

    explicit ElMag3D(Param& p)
        : u(p.x_sampl, p.y_sampl, p.z_sampl, p.dx, p.dy, p.dz), rho(p.x_sampl, p.y_sampl, p.z_sampl, p.dx, p.dy, p.dz),
          voltage(p.x_sampl, p.y_sampl, p.z_sampl, p.dx, p.dy, p.dz), geometry(p)
    {
        mag2d_grid_desc g = {};
        g.coord = MAG2D_CARTESIAN3D; g.boundary = p.boundary; g.mover = MAG2D_ADVANCE_BORIS;
        g.M = p.x_sampl; g.N = p.z_sampl; g.K = p.y_sampl;
        g.x_max = p.x_max; g.z_max = p.z_max; g.y_max = p.y_max;
        g.dx = p.dx; g.dz = p.dz; g.dy = p.dy; g.idx = p.idx; g.idz = p.idz; g.idy = p.idy;
        g.selfconsistent = p.selfconsistent; g.geometry_empty = 0;
        g.magnetic_field_const = 1; g.Br = p.Br; g.Bz = p.Bz; g.Bt = p.Bt; g.dV = p.dV; g.macroparticle_factor = p.macroparticle_factor;
        check(mag2d_create(0, &g, nullptr, &gpu));
        voltage.reset();
        geometry.set_voltage(voltage);
        check(mag2d_set_grid(gpu, reinterpret_cast<const uint8_t*>(geometry.mask.data()), voltage.data()));
    }
    ~ElMag3D() { if (gpu) mag2d_destroy(gpu); }
    // ElMag3D::solve (fields3d.cpp:170-181) -> Solver::solve
    void solve()
    {
        check(mag2d_solve(gpu, 0, 0.0, 0, nullptr, nullptr));
        check(mag2d_get_potential(gpu, 0, u.data()));
    }
    void E(double x, double y, double z, double& Ex, double& Ey, double& Ez) { check(mag2d_field_E3(gpu, 1, &x, &y, &z, &Ex, &Ey, &Ez)); }
    static void check(int rc) { if (rc) throw std::runtime_error(mag2d_last_error()); }
};
