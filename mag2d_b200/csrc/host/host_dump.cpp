// host_dump — prints what the C++ host layer parsed (Param, species, geometry mask) without touching a GPU,
// so that CPU-only tests can compare it field by field with the reference and with the Python readers.
//   host_dump config=... species_conf=...
#include <iomanip>
#include <iostream>

#include "field2d.hpp"
#include "parser.hpp"

// t_grid needs the CUDA library header only for Fields; include the geometry class alone
#define gpu_check mag2d_host_dump_unused
#include "fields.hpp"

int main(int argc, char* argv[])
{
    try
    {
        GetPot cl(argc, argv);
        const std::string field_file = cl("field2d", "");
        if (!field_file.empty())
        {
            // Field2D::load (src/Field2D.cpp:46-130) alone: dimensions, origin, far corner and the values, row by row
            Field2D F(2, 2, 1.0, 1.0);
            F.load(field_file.c_str());
            std::cout << std::setprecision(17) << "field2d " << F.jmax << " " << F.lmax << " " << F.GetXMin() << " " << F.GetYMin() << " " << F.GetXMax() << " "
                      << F.GetYMax() << "\nvalues";
            for (int i = 0; i < F.jmax; i++)
                for (int j = 0; j < F.lmax; j++) std::cout << " " << F[i][j];
            std::cout << "\n";
            return 0;
        }
        GetPot config(cl("config", "config.txt").c_str());
        Param p(config);
        std::cout << std::setprecision(17);
        std::cout << "x_max " << p.x_max << "\nz_max " << p.z_max << "\nx_sampl " << p.x_sampl << "\nz_sampl " << p.z_sampl << "\ndx " << p.dx
                  << "\ndz " << p.dz << "\nidx " << p.idx << "\nidz " << p.idz << "\nV " << p.V << "\ndV " << p.dV << "\ndy " << p.dy
                  << "\nextern_field " << p.extern_field << "\nrf_omega " << p.rf_omega << "\nrf_amplitude " << p.rf_amplitude << "\nniter " << p.niter
                  << "\nmover " << p.mover << "\ncoord " << p.coord << "\nboundary " << p.boundary << "\ngeometry " << p.geometry
                  << "\nselfconsistent " << p.selfconsistent << "\nrf " << p.rf << "\nt_dist_sample " << p.t_dist_sample << "\nt_equilib " << p.t_equilib
                  << "\nmacroparticle_factor " << p.macroparticle_factor << "\nneutral_density " << p.neutral_density << "\n";
        std::vector<SpeciesParams*> vs;
        std::vector<InteractionParams*> vi;
        config_parse(cl("species_conf", "species_conf.txt"), vs, vi);
        for (auto s : vs)
            std::cout << "species " << s->name << " " << s->type << " " << s->mass << " " << s->charge << " " << s->density << " " << s->temperature << " "
                      << s->E_max << " " << s->dt << "\n";
        for (auto q : vi)
            std::cout << "interaction " << q->name << " " << q->type << " " << q->primary << " " << q->secondary << " " << q->DE << " " << q->rate << " "
                      << q->cutoff << " " << q->CS_energy.size() << "\n";
        t_grid grid(p);
        std::cout << "mask";
        for (int i = 0; i < grid.M; i++)
            for (int j = 0; j < grid.N; j++) std::cout << " " << (int)grid.mask[i][j];
        std::cout << "\nvoltage";
        for (int i = 0; i < grid.M; i++)
            for (int j = 0; j < grid.N; j++) std::cout << " " << (grid.mask[i][j] < 2 ? grid.voltage[i][j] : 0.0);
        std::cout << "\n";
        return 0;
    }
    catch (std::exception& e)
    {
        std::cerr << e.what();
        return 1;
    }
}
