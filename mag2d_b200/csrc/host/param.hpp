// param.hpp — Param: the run parameters of config.txt (drop-in for reference src/param.hpp + src/param.cpp).
// Same public fields, enums, key names, defaults, derived quantities and exception texts.
#pragma once
#include <cmath>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>

#include "getpot_lite.hpp"

enum Coord { CARTESIAN, CYLINDRICAL, CARTESIAN3D };

namespace physconst {
// old CODATA values, kept because the reference's results depend on them (src/param.cpp:8-10)
constexpr double eps_0 = 8.854187817e-12;
constexpr double k_B = 1.380662e-23;
constexpr double q_e = 1.602189e-19;
}  // namespace physconst

template <class T>
inline T string2(const std::string& text)
{
    std::istringstream in(text);
    T v;
    if (!(in >> v)) throw std::runtime_error("string2: error converting string " + text + " to T\n");
    return v;
}

class Param
{
  public:
    enum Boundary { FREE, PERIODIC, MIRROR };
    enum Mover { ADVANCE_BORIS, ADVANCE_MULTICOLL, ADVANCE_LEAPFROG };
    enum Geometry { EMPTY, PROBE, RF_22PT, RF_8PT, RF_HAITRAP, RF_QUAD, MAC, PENNING, PENNING_SIMPLE, TUBE };

    double x_max, y_max, z_max;
    double x_min = 0, y_min = 0, z_min = 0;
    int x_sampl, y_sampl, z_sampl;
    double extern_field;
    std::string electric_field_static_file, electric_field_rf_file;
    bool electric_field_from_file;
    std::string magnetic_field_file;
    bool magnetic_field_const;
    double Br, Bz, Bt;
    bool has_probe;
    double probe_radius, probe_length, u_probe;
    double n_particles_total, density_total;
    double dx, dy, dz, V, dV, idx, idy, idz;
    const double eps_0 = physconst::eps_0, k_B = physconst::k_B, q_e = physconst::q_e;
    double pressure, neutral_temperature, macroparticle_factor, dt_elon;
    unsigned long niter;
    Mover mover;
    Coord coord;
    Boundary boundary;
    Geometry geometry;
    int src_fact;
    bool selfconsistent, use_source, u_smooth, rf;
    double rf_amplitude, rf_U0, rf_omega;
    bool particle_reload;
    unsigned long t_print, t_print_dist, t_dist_sample, t_equilib;
    std::string particle_reload_dir = ".", output_dir, species_conf_file;
    bool do_plot;
    double neutral_density;

    explicit Param(GetPot& cfg)
    {
        x_max = cfg("x_max", cfg("r_max", 1e-2));            // r_* are the older spellings
        y_max = cfg("y_max", 1e-2);
        z_max = cfg("z_max", 1e-2);
        x_sampl = cfg("x_sampl", cfg("r_sampl", 100));
        y_sampl = cfg("y_sampl", 2);
        z_sampl = cfg("z_sampl", 100);
        n_particles_total = cfg("n_particles_total", 1e5);
        density_total = cfg("density_total", 1e11);
        pressure = cfg("pressure", 133.0);
        neutral_temperature = cfg("neutral_temperature", 300.0);
        probe_radius = cfg("probe_radius", 1e-4);
        probe_length = cfg("probe_length", 1e-2);
        u_probe = cfg("u_probe", -10.0);
        extern_field = cfg("extern_field", 100.0);
        electric_field_static_file = cfg("electric_field_static_file", "");
        electric_field_rf_file = cfg("electric_field_rf_file", "");
        electric_field_from_file = cfg("electric_field_from_file", 0) != 0;
        has_probe = cfg("has_probe", 0) != 0;
        magnetic_field_file = cfg("magnetic_field_file", "");
        magnetic_field_const = cfg("magnetic_field_const", 1) != 0;
        Br = cfg("Br", 0.0);
        Bz = cfg("Bz", 0.0);
        Bt = cfg("Bt", 0.0);
        niter = string2<unsigned long>(cfg("niter", "100000"));
        dt_elon = cfg("dt_elon", 1e-11);
        selfconsistent = cfg("selfconsistent", 1) != 0;
        use_source = cfg("use_source", 0) != 0;
        u_smooth = cfg("u_smooth", 0) != 0;
        rf = cfg("rf", 1) != 0;
        rf_amplitude = cfg("rf_amplitude", 10.0);
        rf_U0 = cfg("rf_U0", 0.0);
        rf_omega = cfg("rf_omega", 2 * M_PI * cfg("rf_freq", 20e6));
        if (selfconsistent && rf) throw std::runtime_error("Param: selfconsistent rf trap not implemented\n");
        if (selfconsistent && electric_field_from_file)
            throw std::runtime_error("Param: selfconsistent with electric_field_from_file not implemented");
        t_print = string2<unsigned long>(cfg("t_print", "0"));
        t_print_dist = string2<unsigned long>(cfg("t_print_dist", "0"));
        t_dist_sample = t_print > 10 ? t_print / 10 : 1;
        const std::string eq = cfg("t_equilib", "niter+1");
        t_equilib = eq == "niter+1" ? niter + 1 : string2<unsigned long>(eq);
        particle_reload = cfg("particle_reload", 0) != 0;
        particle_reload_dir = cfg("particle_reload_dir", ".");
        src_fact = cfg("src_fact", 20);
        neutral_density = pressure / (k_B * neutral_temperature);
        do_plot = cfg("do_plot", 1) != 0;

        coord = pick<Coord>(cfg("coord", "CYLINDRICAL"), {{"CYLINDRICAL", CYLINDRICAL}, {"CARTESIAN", CARTESIAN}, {"CARTESIAN3D", CARTESIAN3D}}, "coord");
        boundary = pick<Boundary>(cfg("boundary", "FREE"), {{"FREE", FREE}, {"MIRROR", MIRROR}, {"PERIODIC", PERIODIC}}, "boundary");
        if (coord == CYLINDRICAL && boundary != FREE)
            throw std::runtime_error("Param: only FREE boundary condition in cylindrical coords is implemented\n");
        if (boundary == MIRROR) throw std::runtime_error("Param: MIRROR boundary condition not implemented\n");
        // the leapfrog mover exists in the reference but cannot be selected (param.cpp:98-101)
        mover = pick<Mover>(cfg("mover", "ADVANCE_BORIS"), {{"ADVANCE_BORIS", ADVANCE_BORIS}, {"ADVANCE_MULTICOLL", ADVANCE_MULTICOLL}}, "mover");
        geometry = pick<Geometry>(cfg("geometry", "EMPTY"),
                                  {{"EMPTY", EMPTY}, {"PROBE", PROBE}, {"RF_22PT", RF_22PT}, {"RF_8PT", RF_8PT}, {"RF_HAITRAP", RF_HAITRAP},
                                   {"RF_QUAD", RF_QUAD}, {"MAC", MAC}, {"PENNING", PENNING}, {"PENNING_SIMPLE", PENNING_SIMPLE}, {"TUBE", TUBE}},
                                  "geometry");
        dx = x_max / (x_sampl - 1);
        dz = z_max / (z_sampl - 1);
        idx = 1.0 / dx;
        idz = 1.0 / dz;
        cfg.set_prefix("");
        macroparticle_factor = cfg("macroparticle_factor", 1e4);
        V = n_particles_total / density_total;            // the cell depth follows from the particle count
        dV = V / ((x_sampl - 1) * (z_sampl - 1));
        dy = dV / (dx * dz);
        if (coord == CYLINDRICAL) dy = 2 * M_PI / macroparticle_factor;
        idy = 1.0 / dy;
    }

  private:
    template <class E>
    static E pick(const std::string& word, const std::map<std::string, E>& table, const char* what)
    {
        auto it = table.find(word);
        if (it == table.end()) throw std::runtime_error(std::string("Param: unrecognized ") + what + " value " + word + "\n");
        return it->second;
    }
};
