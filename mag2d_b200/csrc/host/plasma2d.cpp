// plasma2d — the reference's main driver (src/test.cpp) on top of libmag2d_b200.
//   plasma2d_b200 config=config.txt species_conf=species_conf.txt initscript=initscript.txt output_dir=output
// coord = CARTESIAN or CYLINDRICAL is taken from the config (the reference binary is Cartesian-only because
// its cylindrical loaders are not reachable from the initscript language).
#include <fstream>
#include <iostream>

#include "output.hpp"
#include "pic.hpp"

template <int D>
static int run(Param& param, const std::string& initscript)
{
    Pic<D> pic(param);
    pic.check_params();
    pic.run_initscript(initscript);
    pic.advance_init();
    if (param.use_source)      // test.cpp:56-59
        for (size_t i = 0; i < pic.speclist.size(); i++)
            if (pic.speclist[i]->particle) pic.speclist[i]->source5_refresh(param.src_fact);
    std::ofstream fw((param.output_dir + "/out.dat").c_str());
    for (unsigned long i = 1; i < param.niter + 1; ++i)
    {
        pic.advance();
        if (i % 10 == 0) pic.print_trace();
        if (i % param.t_dist_sample == 0) pic.dist_sample();
        if ((param.t_print_dist != 0 && i % param.t_print_dist == 0) || i == param.niter)
        {
            pic.print_distribution();
            pic.print_field();
        }
        if (param.t_print != 0 && i % param.t_print == 0)
        {
            pic.print_status(fw);
            std::cout << "plot " << i << std::endl;
            if (i < param.t_equilib) pic.dist_reset();
        }
    }
    return 0;
}

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
        if (param.coord == CYLINDRICAL) return run<CYLINDRICAL>(param, initscript);
        return run<CARTESIAN>(param, initscript);
    }
    catch (std::exception& e)
    {
        std::cerr << "terminate called after throwing an instance of 'std::runtime_error'\n  what():  " << e.what() << std::endl;
        return 134;
    }
}
