// parser.hpp — species_conf.txt reader (drop-in for reference src/parser.hpp + src/parser.cpp): SPECIES /
// INTERACTION / CROSS_SECTION ... END_CROSS_SECTION blocks of "KEY value" lines, '#' comment lines.
// Unlike the reference (parser.cpp:45 "TODO specify defaults") every field starts at zero instead of heap garbage.
#pragma once
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "param.hpp"

enum SpeciesType { NEUTRAL, ELECTRON, ION };
enum CollType { ELASTIC, LANGEVIN, CX, COULOMB, SUPERELASTIC };

struct SpeciesParams
{
    std::string name;
    SpeciesType type = NEUTRAL;
    double charge = 0, mass = 0, dt = 0, density = 0, temperature = 0, polarizability = 0, E_max = 0;
};

struct InteractionParams
{
    std::string name;
    CollType type = ELASTIC;
    double DE = 0, rate = 0, cutoff = 0;
    std::string primary, secondary;
    std::vector<double> CS_energy, CS_value;
};

inline void config_parse(const std::string& fname, std::vector<SpeciesParams*>& species, std::vector<InteractionParams*>& interactions)
{
    std::ifstream in(fname.c_str());
    enum { NONE, DEFAULTS, IN_SPECIES, IN_INTERACTION, IN_TABLE } where = NONE;
    std::string line;
    while (std::getline(in, line))
    {
        if (!line.empty() && line[0] == '#') continue;
        std::istringstream words(line);
        std::vector<std::string> w;
        for (std::string t; words >> t;) w.push_back(t);
        if (w.empty()) continue;
        const std::string& key = w[0];
        if (key == "DEFAULT") { where = DEFAULTS; continue; }
        if (key == "SPECIES") { species.push_back(new SpeciesParams); where = IN_SPECIES; continue; }
        if (key == "INTERACTION") { interactions.push_back(new InteractionParams); where = IN_INTERACTION; continue; }
        if (where == NONE) throw std::runtime_error("config_parse: unrecognized first config block\n");
        const std::string arg = w.size() > 1 ? w[1] : std::string();
        if (where == IN_SPECIES)
        {
            SpeciesParams& s = *species.back();
            if (key == "NAME") s.name = arg;
            else if (key == "TYPE")
            {
                if (arg == "NEUTRAL") s.type = NEUTRAL;
                else if (arg == "ELECTRON") s.type = ELECTRON;
                else if (arg == "ION") s.type = ION;
                else throw std::runtime_error("config_parse: unrecognized first species type \"" + arg + "\"");
            }
            else if (key == "MASS") s.mass = string2<double>(arg);
            else if (key == "CHARGE") s.charge = string2<double>(arg);
            else if (key == "DENSITY") s.density = string2<double>(arg);
            else if (key == "DT") s.dt = string2<double>(arg);
            else if (key == "TEMPERATURE") s.temperature = string2<double>(arg);
            else if (key == "EMAX") s.E_max = string2<double>(arg);
            else throw std::runtime_error("config_parse: unrecognized species  parameter \"" + key + "\"");
        }
        else if (where == IN_INTERACTION)
        {
            InteractionParams& q = *interactions.back();
            if (key == "NAME") q.name = arg;
            else if (key == "TYPE")
            {
                if (arg == "ELASTIC") q.type = ELASTIC;
                else if (arg == "LANGEVIN") q.type = LANGEVIN;
                else if (arg == "CX") q.type = CX;
                else if (arg == "COULOMB") q.type = COULOMB;
                else if (arg == "SUPERELASTIC") q.type = SUPERELASTIC;
                else throw std::runtime_error("config_parse: unrecognized interaction type \"" + arg + "\"");
            }
            else if (key == "DE") q.DE = string2<double>(arg);
            else if (key == "RATE") q.rate = string2<double>(arg);
            else if (key == "CUTOFF") q.cutoff = string2<double>(arg);
            else if (key == "PRIMARY") q.primary = arg;
            else if (key == "SECONDARY") q.secondary = arg;
            else if (key == "CROSS_SECTION") where = IN_TABLE;
            else throw std::runtime_error("config_parse: unrecognized species  parameter\"" + key + "\"");
        }
        else if (where == IN_TABLE)
        {
            if (key == "END_CROSS_SECTION") where = IN_INTERACTION;
            else
            {
                interactions.back()->CS_energy.push_back(string2<double>(key));
                interactions.back()->CS_value.push_back(string2<double>(arg));
            }
        }
    }
}
