// checkpoint_b200 — the reference's restart path either side of the device store, without stepping:
//   checkpoint_b200 config=config.txt species_conf=species_conf.txt output_dir=out
// Pic's constructor reads particles_<NAME>.dat (+ particles_source_<NAME>.dat) from particle_reload_dir when
// particle_reload = 1 (src/pic.cpp:136-145 -> BaseSpecies::load, src/particles.cpp:61-93) and uploads them; Pic::save
// (src/pic.cpp:462-470 -> BaseSpecies::save, src/particles.cpp:32-59) downloads them again and writes the same files
// into output_dir.  The reference's drivers hard-wire t_save = 0 (src/test.cpp:16), so its own save is only reachable
// from code; tests/test_reference_files.py drives both sides.
#include <iostream>

#include "output.hpp"
#include "pic.hpp"

template <int D>
static int run(Param& param)
{
    Pic<D> pic(param);
    pic.save();
    for (size_t i = 0; i < pic.speclist.size(); i++)
        if (pic.speclist[i]->particle) std::cout << pic.speclist[i]->name << " " << pic.speclist[i]->n_particles() << std::endl;
    return 0;
}

int main(int argc, char* argv[])
{
    try
    {
        GetPot cl(argc, argv);
        GetPot config(cl("config", "config.txt").c_str());
        Param param(config);
        param.species_conf_file = cl("species_conf", "species_conf.txt");
        param.output_dir = cl("output_dir", "output");
        t_output output(param.output_dir);
        if (param.coord == CYLINDRICAL) return run<CYLINDRICAL>(param);
        return run<CARTESIAN>(param);
    }
    catch (std::exception& e)
    {
        std::cerr << e.what() << std::endl;
        return 134;
    }
}
