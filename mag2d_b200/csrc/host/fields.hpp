// fields.hpp — t_grid (electrode geometry) and Fields (potential / charge grids) host classes: drop-in for
// reference src/fields.hpp + src/fields.cpp with the UMFPACK factorisation and solves replaced by the
// multigrid solver of libmag2d_b200 (mag2d_solve).  u, uRF, rho are host mirrors of device arrays:
// download() refreshes them for printing, upload() pushes edits back.
#pragma once
#include <cmath>
#include <stdexcept>
#include <string>

#include "../../../include/mag2d_b200.h"
#include "field2d.hpp"
#include "param.hpp"

enum { FIXED, FIXED_RF, FREE, BOUNDARY };   // node classes, values as in the reference (fields.hpp:20)

inline void gpu_check(int rc)
{
    if (rc) throw std::runtime_error(mag2d_last_error());
}

class t_grid
{
  public:
    int M, N;
    double dx, dz;
    Array2D<char> mask;
    Array2D<double> voltage;
    double U_trap = 0;
    Param* p_param;

    explicit t_grid(Param& param) : M(param.x_sampl), N(param.z_sampl), dx(param.dx), dz(param.dz), mask(M, N), voltage(M, N), p_param(&param)
    {
        switch (param.geometry)
        {
            case Param::PROBE: probe(); break;
            case Param::PENNING_SIMPLE: penning_trap_simple(); break;
            case Param::PENNING: penning_trap(); break;
            case Param::MAC: MAC_filter(); break;
            case Param::RF_QUAD: rf_trap(); break;
            case Param::RF_HAITRAP: rf_multipole(8, 0.3e-2 + 0.01e-2, 0.01e-2, true); break;
            case Param::RF_8PT: rf_multipole(8, 0.3e-2 + 0.1e-2, 0.1e-2, true); break;
            case Param::RF_22PT: rf_multipole(22, 0.75e-2, 0.05e-2, false); break;
            case Param::TUBE: tube(); break;
            default: empty(); break;
        }
    }
    // a particle survives while any corner of its cell is a FREE node (fields.hpp:94-101)
    bool is_free(double r, double z) const
    {
        const int i = (int)(r * p_param->idx), j = (int)(z * p_param->idz);
        return mask[i][j] == FREE || mask[i + 1][j] == FREE || mask[i][j + 1] == FREE || mask[i + 1][j + 1] == FREE;
    }
    void square_electrode(double rmin, double rmax, double zmin, double zmax, double v, char kind = FIXED)
    {
        for (int i = 0; i < M; i++)
            for (int j = 0; j < N; j++)
                if (i * dx > rmin && i * dx < rmax && j * dz > zmin && j * dz < zmax) { mask[i][j] = kind; voltage[i][j] = v; }
    }
    void circle_electrode(double rc, double zc, double radius, double v, char kind = FIXED)
    {
        for (int i = 0; i < M; i++)
            for (int j = 0; j < N; j++)
                if ((i * dx - rc) * (i * dx - rc) + (j * dz - zc) * (j * dz - zc) <= radius * radius) { mask[i][j] = kind; voltage[i][j] = v; }
    }
    void empty() { frame(false, true); }
    void probe()
    {
        frame(false, true);
        circle_electrode((M - 1) * dx / 2, (N - 1) * dz / 2, p_param->probe_radius, p_param->u_probe, FIXED);
        mark_boundary();
    }
    void tube()
    {
        frame(false, false);
        const double c = p_param->x_max / 2.0, r2 = p_param->probe_radius * p_param->probe_radius;
        for (int i = 0; i < M; i++)
            for (int j = 0; j < N; j++)
                if ((i * dx - c) * (i * dx - c) + (j * dz - c) * (j * dz - c) >= r2) { mask[i][j] = FIXED; voltage[i][j] = 0.0; }
        mark_boundary();
    }
    void rf_trap()
    {
        frame(false, false);
        circle_electrode(5e-3, 1e-2, 2e-3, 1.0, FIXED_RF);
        circle_electrode(15e-3, 1e-2, 2e-3, 1.0, FIXED_RF);
        circle_electrode(1e-2, 5e-3, 2e-3, -1.0, FIXED_RF);
        circle_electrode(1e-2, 15e-3, 2e-3, -1.0, FIXED_RF);
        mark_boundary();
    }
    // npoles rods of alternating RF polarity on a ring around (1 cm, 1 cm): the 22-pole, 8-pole and "haitrap" geometries
    void rf_multipole(int npoles, double r_ring, double r_rod, bool ramp)
    {
        frame(false, ramp);
        for (int k = 0; k < npoles; k++)
        {
            const double a = 2 * M_PI * (k + 1.0 / 32) / npoles;
            circle_electrode(1e-2 + sin(a) * r_ring, 1e-2 + cos(a) * r_ring, r_rod, k % 2 == 0 ? -1 : 1, FIXED_RF);
        }
        mark_boundary();
    }
    void rf_22PT() { rf_multipole(22, 0.75e-2, 0.05e-2, false); }
    void rf_8PT() { rf_multipole(8, 0.3e-2 + 0.1e-2, 0.1e-2, true); }
    void rf_haitrap() { rf_multipole(8, 0.3e-2 + 0.01e-2, 0.01e-2, true); }
    void MAC_filter()
    {
        frame(true, false);
        const double th = p_param->u_probe, ofs = 3e-2;
        square_electrode(5e-3, 4.5e-2, 1e-2, 1.5e-2, -.00);
        square_electrode(5e-3, 7e-3, 2e-2, 8e-2, 0.0);
        square_electrode(5e-3, 4.5e-2, 8.5e-2, 9e-2, -.00);
        square_electrode(3e-2, 3.3e-2, 11e-2 + ofs, 14e-2 + ofs, 0.8 * th);
        square_electrode(4.5e-2, 4.8e-2, 15e-2 + ofs, 25e-2 + ofs, th);
        square_electrode(3e-2, 3.3e-2, 26e-2 + ofs, 29e-2 + ofs, 1.0 * th);
        square_electrode(2.5e-2, 2.8e-2, 29e-2 + ofs, 30.5e-2 + ofs, 1.0 * th);
        square_electrode(15e-3, 4.5e-2, 35e-2, 35.3e-2, .0);
        square_electrode(0.0, 4.5e-2, 39.5e-2, 40e-2, 3e3);
        mark_boundary();
    }
    void penning_trap()
    {
        frame(true, false);
        square_electrode(1.57e-2 / 2, 1.67e-2 / 2, 1e-3, 25e-3, -5);
        square_electrode(0, 1.46e-2 / 2, 15e-3, 16e-3, 10);
        square_electrode(0, 7e-3 / 2, 12e-3, 19e-3, 10);
        square_electrode(0, 1.9e-3, 1e-3, 12e-3, 10);
        square_electrode(4e-3, 7e-3, 52e-3, 53e-3, -5);
        square_electrode(2.5e-3, 7e-3, 46e-3, 47e-3, 0);
        mark_boundary();
    }
    void penning_trap_simple(double trap_voltage = -1.0)
    {
        U_trap = trap_voltage;
        frame(true, false);
        const double ri = 1e-2, ro = 1.1e-2;
        square_electrode(ri, ro, 0, 1e-2, -0.5);
        square_electrode(ri, ro, 1.1e-2, 2e-2, 0);
        square_electrode(ri, ro, 2.1e-2, 6e-2, trap_voltage);
        square_electrode(ri, ro, 6.1e-2, 7.5e-2, -10);
        mark_boundary();
    }

  private:
    // Dirichlet frame: all four edges, or three when the r = 0 axis stays open; optional linear ramp that
    // carries extern_field along z
    void frame(bool axis_open, bool ramp)
    {
        for (int i = 0; i < M; i++)
            for (int j = 0; j < N; j++)
            {
                const bool edge = i == M - 1 || j == 0 || j == N - 1 || (!axis_open && i == 0);
                mask[i][j] = edge ? FIXED : FREE;
                voltage[i][j] = (edge && ramp) ? -p_param->extern_field * dz * (j - N / 2) : 0.0;
            }
    }
    void mark_boundary()
    {
        for (int i = 2; i < M - 2; i++)
            for (int j = 2; j < N - 2; j++)
                if (mask[i][j] != FIXED && (mask[i - 1][j] == FIXED || mask[i + 1][j] == FIXED || mask[i][j - 1] == FIXED || mask[i][j + 1] == FIXED))
                    mask[i][j] = BOUNDARY;
    }
};

class Fields
{
  public:
    t_grid grid;
    Field2D u, uRF, uTmp, uAvg, rho;
    mag2d_ctx* gpu = nullptr;

    explicit Fields(Param& param)
        : grid(param), u(param.x_sampl, param.z_sampl, param.dx, param.dz), uRF(param.x_sampl, param.z_sampl, param.dx, param.dz),
          uTmp(param.x_sampl, param.z_sampl, param.dx, param.dz), uAvg(param.x_sampl, param.z_sampl, param.dx, param.dz),
          rho(param.x_sampl, param.z_sampl, param.dx, param.dz), p_param(&param)
    {
        mag2d_grid_desc g = {};
        g.coord = param.coord; g.boundary = param.boundary; g.mover = param.mover;
        g.M = param.x_sampl; g.N = param.z_sampl; g.K = param.y_sampl;
        g.x_max = param.x_max; g.z_max = param.z_max; g.y_max = param.y_max;
        g.dx = param.dx; g.dz = param.dz; g.dy = param.dy; g.idx = param.idx; g.idz = param.idz; g.idy = param.idy;
        g.selfconsistent = param.selfconsistent; g.rf = param.rf; g.geometry_empty = param.geometry == Param::EMPTY;
        g.electric_field_from_file = param.electric_field_from_file; g.extern_field = param.extern_field;
        g.rf_amplitude = param.rf_amplitude; g.rf_U0 = param.rf_U0; g.rf_omega = param.rf_omega;
        g.magnetic_field_const = param.magnetic_field_const; g.u_smooth = param.u_smooth;
        g.Br = param.Br; g.Bz = param.Bz; g.Bt = param.Bt; g.dV = param.dV; g.macroparticle_factor = param.macroparticle_factor;
        gpu_check(mag2d_create(0, &g, nullptr, &gpu));
        gpu_check(mag2d_set_use_source(gpu, param.use_source));
        gpu_check(mag2d_set_grid(gpu, reinterpret_cast<const uint8_t*>(grid.mask[0]), grid.voltage[0]));
    }
    ~Fields() { mag2d_destroy(gpu); }
    Fields(const Fields&) = delete;

    // replace the umfpack_di_solve calls of fields.cpp:311,347,352
    void boundary_solve() { gpu_check(mag2d_solve(gpu, 0, solve_tol, 100, nullptr, nullptr)); }
    void boundary_solve_rf() { gpu_check(mag2d_solve(gpu, 1, solve_tol, 100, nullptr, nullptr)); }
    void solve() { boundary_solve(); }
    void reset() { gpu_check(mag2d_rho_reset(gpu, -1)); rho.reset(); }
    void u_smooth(bool symmetry = false, double radius = -1.) { gpu_check(mag2d_u_smooth(gpu, symmetry, radius)); }
    void download()
    {
        gpu_check(mag2d_get_potential(gpu, 0, u[0]));
        gpu_check(mag2d_get_potential(gpu, 1, uRF[0]));
    }
    void upload()
    {
        gpu_check(mag2d_set_potential(gpu, 0, u[0]));
        gpu_check(mag2d_set_potential(gpu, 1, uRF[0]));
    }
    // E = -grad(u + uRF*(A cos(wt) + U0)) or the constant external field (fields.hpp:124-150); host diagnostics
    void E(double x, double y, double& ex, double& ez, double time = 0) const
    {
        if (p_param->geometry == Param::EMPTY && !p_param->selfconsistent) { ex = 0.; ez = p_param->extern_field; return; }
        gpu_check(mag2d_field_E(gpu, 1, &x, &y, time, &ex, &ez));
    }
    // fields.hpp:152-177: the constants of config.txt or the interpolated table (host diagnostics; the movers read the
    // device copy of the table)
    void B(double x, double y, double& br, double& bz, double& bt) const
    {
        if (p_param->magnetic_field_const) { br = p_param->Br; bz = p_param->Bz; bt = p_param->Bt; return; }
        br = Br.interpolate(x, y);
        bz = Bz.interpolate(x, y);
        bt = 0.00;
    }
    // fields.cpp:870-959: four-column file "r z Br Bz" on a regular grid, either row order; fills Br / Bz and hands the
    // table to the GPU context
    void load_magnetic_field(const char* fname)
    {
        std::ifstream fr(fname);
        if (fr.fail()) throw std::runtime_error("Fields::load_magnetic_field(): failed opening file\n");
        std::vector<double> rvec, zvec, brvec, bzvec;
        for (std::string line; std::getline(fr, line);)
        {
            std::istringstream row(line);
            double a, b, c, d;
            if (row >> a >> b >> c >> d) { rvec.push_back(a); zvec.push_back(b); brvec.push_back(c); bzvec.push_back(d); }
        }
        if (rvec.size() < 2) throw std::runtime_error("Fields::load_magnetic_field() wrong size of input vector");
        auto to_int = [](double x) {       // double2int, util.cpp:22-28
            const int res = (int)(x + 0.5);
            if (std::fabs(res - x) > 1e-2) throw std::runtime_error("double2int() " + std::to_string(x) + " is not integer\n");
            return res;
        };
        auto axis = [&](const std::vector<double>& v, double& lo, double& step) {
            lo = v.front();
            double hi = v.back();
            size_t k = 1;
            while (k < v.size() && v[k] - v[k - 1] == 0.0) k++;
            if (k >= v.size()) throw std::runtime_error("Fields::load_magnetic_field() wrong size of input vector");
            step = v[k] - v[k - 1];
            if (step < 0) { step = -step; lo = v.back(); hi = v.front(); }
            return (unsigned)to_int((hi - lo) / step + 1);
        };
        double rmin, zmin, dr, dz;
        const unsigned rsampl = axis(rvec, rmin, dr), zsampl = axis(zvec, zmin, dz);
        if ((size_t)rsampl * zsampl != rvec.size()) throw std::runtime_error("Fields::load_magnetic_field() wrong size of input vector");
        Br.resize(rsampl, zsampl, dr, dz, rmin, zmin);
        Bz.resize(rsampl, zsampl, dr, dz, rmin, zmin);
        for (unsigned i = 0; i < rsampl; i++)
            for (unsigned j = 0; j < zsampl; j++) Br[i][j] = Bz[i][j] = std::numeric_limits<double>::quiet_NaN();
        for (size_t k = 0; k < rvec.size(); k++)
        {
            const int ri = to_int((rvec[k] - rmin) / dr), zi = to_int((zvec[k] - zmin) / dz);
            Br[ri][zi] = brvec[k];
            Bz[ri][zi] = bzvec[k];
        }
        if (Br.hasnan() || Bz.hasnan()) throw std::runtime_error("Fields::load_magnetic_field() garbage loaded");
        gpu_check(mag2d_set_magnetic_field(gpu, (int)rsampl, (int)zsampl, dr, dz, rmin, zmin, Br[0], Bz[0]));
    }
    Field2D Br, Bz;     // private in the reference (fields.hpp:78); public here for the dump tool
    void u_sample() { download(); uAvg.add(u); nsampl++; }
    void u_reset() { uAvg.reset(); nsampl = 0; }
    void u_print(const char* fname) { uAvg.print(fname, 1.0 / nsampl); }
    double solve_tol = 1e-13;

  private:
    Param* p_param;
    int nsampl = 0;
};
