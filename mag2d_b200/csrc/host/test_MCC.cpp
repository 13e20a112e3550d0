// test_MCC — the reference's collision benchmark driver (src/test_MCC.cpp) on top of libmag2d_b200: every
// iteration removes all particles, re-runs the initscript and times one Pic::advance.
#include <fstream>
#include <iostream>

#include "output.hpp"
#include "pic.hpp"

int main(int argc, char* argv[])
{
    try
    {
        GetPot cl(argc, argv);
        const std::string config_file = cl("config", "config.txt");
        const std::string species_conf_file = cl("species_conf", "species_conf.txt");
        const std::string initscript = cl("initscript", "initscript.txt");
        GetPot config(config_file.c_str());
        Param param(config);
        param.species_conf_file = species_conf_file;
        param.output_dir = cl("output_dir", "output");
        t_output output(param.output_dir);
        output.backup(config_file, "config.txt");
        output.backup(species_conf_file, "species_conf.txt");
        output.backup(initscript, "initscript.txt");

        Pic<CARTESIAN> pic(param);
        gpu_check(mag2d_seed(pic.field.gpu, 1234));      // pic.rnd.initialize_seed(1234)
        pic.run_initscript(initscript);
        std::ofstream fw((param.output_dir + "/out.dat").c_str());
        double seconds = 0;
        for (unsigned long i = 1; i < param.niter + 1; ++i)
        {
            for (auto s : pic.speclist) s->remove_all();
            pic.run_initscript(initscript);
            const double before = pic.step_seconds();
            pic.advance();
            seconds += pic.step_seconds() - before;
            if ((param.t_print_dist != 0 && i % param.t_print_dist == 0) || i == param.niter)
            {
                pic.dist_sample();
                pic.print_distribution();
                pic.print_field();
            }
            if (param.t_print != 0 && i % param.t_print == 0)
            {
                pic.print_status(fw);
                std::cout << "plot " << i << " " << seconds / i * 1000 << " ms / iteration" << std::endl;
                if (i < param.t_equilib) pic.dist_reset();
            }
        }
        return 0;
    }
    catch (std::exception& e)
    {
        std::cerr << "terminate called after throwing an instance of 'std::runtime_error'\n  what():  " << e.what() << std::endl;
        return 134;
    }
}
