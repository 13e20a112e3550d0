// plasma3d — the reference's 3-D test driver (src/plasma3d.cpp: does not build there) on top of libmag2d_b200:
// solve the vacuum field of the point electrode, dump it (field3d.dat, field3d.vtk, voltage.dat) and trace one
// 0.03 eV electron for niter steps (traj1.dat, vel1.dat).
//   plasma3d_b200 config=config.txt output_dir=output
#include <cmath>
#include <fstream>
#include <iostream>

#include "fields3d.hpp"
#include "output.hpp"

int main(int argc, char* argv[])
{
    try
    {
        GetPot cl(argc, argv);
        const std::string config_file = cl("config", "config.txt");
        GetPot config(config_file.c_str());
        Param param(config);
        param.output_dir = cl("output_dir", "output");
        t_output output(param.output_dir);

        ElMag3D elmag(param);
        elmag.solve();
        elmag.u.print((param.output_dir + "/field3d.dat").c_str());
        elmag.u.print((param.output_dir + "/field3d.vtk").c_str(), "vtk");
        elmag.voltage.print((param.output_dir + "/voltage.dat").c_str());

        // Elon: one electron (mass 9.109534e-31, charge -1.6e-19) without collisions (plasma3d.cpp:8-16,56-67)
        const double mass = 9.109534e-31, charge = -1.6e-19;
        mag2d_species_desc sd = {MAG2D_ELECTRON, 0, mass, charge, 0.0, 300.0, 1.0, param.dt_elon > 0 ? param.dt_elon : 1e-11};
        ElMag3D::check(mag2d_set_species(elmag.gpu, 1, &sd, 0, nullptr, nullptr, nullptr, 0));
        mag2d_particle p = {};
        p.x = param.x_max * 0.5;
        p.y = param.y_max * 0.5;
        p.z = param.z_max * 0.7;
        p.vx = std::sqrt(0.03 * param.q_e / mass * 2.0);      // veV(0.03)
        ElMag3D::check(mag2d_particles_upload(elmag.gpu, 0, &p, 1));

        std::ofstream fw((param.output_dir + "/traj1.dat").c_str());
        std::ofstream fwv((param.output_dir + "/vel1.dat").c_str());
        for (unsigned long i = 0; i < param.niter; i++)
        {
            int64_t n = 0;
            mag2d_particle q = {};
            ElMag3D::check(mag2d_particles_download(elmag.gpu, 0, &q, 1, &n));
            if (n < 1 || q.empty) break;                      // the particle left the box or hit the electrode
            fw << q.x << " " << q.y << " " << q.z << std::endl;
            fwv << q.vx << " " << q.vy << " " << q.vz << std::endl;
            ElMag3D::check(mag2d_species_advance(elmag.gpu, 0));
        }
        return 0;
    }
    catch (std::exception& e)
    {
        std::cerr << "terminate called after throwing an instance of 'std::runtime_error'\n  what():  " << e.what() << std::endl;
        return 134;
    }
}
