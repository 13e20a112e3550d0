// pic.hpp — Speclist<D> and Pic<D>: the orchestrator surface of reference src/pic.cpp (construction from
// species_conf.txt, initscript interpreter, advance/advance_init, sampling and printing), forwarding the hot
// path to libmag2d_b200.
#pragma once
#include <map>
#include <sys/resource.h>
#include <sys/time.h>

#include "species.hpp"

// user-CPU and wall-clock stopwatch (reference src/timer.hpp)
class t_timer
{
  public:
    void start() { getrusage(RUSAGE_SELF, &u0); gettimeofday(&w0, nullptr); }
    void stop()
    {
        rusage u1;
        timeval w1;
        getrusage(RUSAGE_SELF, &u1);
        gettimeofday(&w1, nullptr);
        cpu += (u1.ru_utime.tv_sec - u0.ru_utime.tv_sec) + 1e-6 * (u1.ru_utime.tv_usec - u0.ru_utime.tv_usec);
        real += (w1.tv_sec - w0.tv_sec) + 1e-6 * (w1.tv_usec - w0.tv_usec);
    }
    double get_cpu_time() { return cpu; }
    double get_real_time() { return real; }
    void reset() { cpu = real = 0; }

  private:
    rusage u0;
    timeval w0;
    double cpu = 0, real = 0;
};

template <int D>
class Speclist
{
  public:
    std::vector<Species<D>*> data;
    Speclist(const std::string& configfile, Param& param, Fields& fields)
    {
        std::vector<SpeciesParams*> vs;
        std::vector<InteractionParams*> vi;
        config_parse(configfile, vs, vi);
        for (size_t i = 0; i < vs.size(); i++) data.push_back(new Species<D>(vs[i], param, fields, (int)i));
        auto index_of = [&](const std::string& who, const std::string& inter) {
            for (size_t k = 0; k < data.size(); k++)
                if (data[k]->name == who) return (int)k;
            throw std::runtime_error("Speclist::Speclist: unrecognized primary species \"" + who + "\" of interaction \"" + inter + "\"\n");
        };
        std::vector<mag2d_species_desc> sd(vs.size());
        for (size_t i = 0; i < vs.size(); i++)
            sd[i] = {(int32_t)vs[i]->type, 0, vs[i]->mass, vs[i]->charge, vs[i]->density, vs[i]->temperature, vs[i]->E_max, vs[i]->dt};
        std::vector<mag2d_interaction_desc> idesc(vi.size());
        std::vector<double> tE, tS;
        for (size_t k = 0; k < vi.size(); k++)
        {
            idesc[k] = {(int32_t)vi[k]->type, index_of(vi[k]->primary, vi[k]->name), index_of(vi[k]->secondary, vi[k]->name),
                        (int32_t)vi[k]->CS_energy.size(), (int32_t)tE.size(), 0, vi[k]->DE, vi[k]->rate, vi[k]->cutoff};
            tE.insert(tE.end(), vi[k]->CS_energy.begin(), vi[k]->CS_energy.end());
            tS.insert(tS.end(), vi[k]->CS_value.begin(), vi[k]->CS_value.end());
        }
        // wires the interaction lists and runs lifetime_init() for every species (particles.cpp:142-170)
        gpu_check(mag2d_set_species(fields.gpu, (int)sd.size(), sd.data(), (int)idesc.size(), idesc.data(), tE.data(), tS.data(), (int)tE.size()));
        for (size_t i = 0; i < data.size(); i++)
        {
            gpu_check(mag2d_species_get(fields.gpu, (int)i, 0, &data[i]->lifetime));
            data[i]->rates_by_species.resize(data.size());
            gpu_check(mag2d_species_rates(fields.gpu, (int)i, data[i]->rates_by_species.data()));
        }
        for (auto p : vs) delete p;
        for (auto p : vi) delete p;
    }
    ~Speclist() { for (auto s : data) delete s; }
    auto begin() { return data.begin(); }
    auto end() { return data.end(); }
    size_t size() { return data.size(); }
    Species<D>* operator[](size_t i) { return data[i]; }
    Species<D>* operator[](const std::string& name)
    {
        for (auto s : data)
            if (s->name == name) return s;
        return nullptr;
    }
};

template <int D>
class Pic
{
  public:
    Param& param;
    std::map<std::string, SpeciesType> string2speciestype;   // species name -> index, as in the reference
    Fields field;
    Speclist<D> speclist;
    unsigned long iter = 0;
    int nsampl = 0;

    explicit Pic(Param& p) : param(p), field(p), speclist(p.species_conf_file, p, field)
    {
        for (size_t i = 0; i < speclist.size(); i++) string2speciestype[speclist[i]->name] = (SpeciesType)i;
        if (param.particle_reload)
            for (size_t i = 0; i < speclist.size(); i++)
                if (speclist[i]->particle)          // pic.cpp:136-145
                {
                    speclist[i]->load(param.particle_reload_dir + "/particles_" + speclist[i]->name + ".dat");
                    speclist[i]->source5_load(param.particle_reload_dir + "/particles_source_" + speclist[i]->name + ".dat", param.src_fact);
                }
        if (!param.magnetic_field_const) field.load_magnetic_field(param.magnetic_field_file.c_str());     // pic.cpp:148-149
        dist_reset();
        if (param.electric_field_from_file)
        {
            field.u.load(param.electric_field_static_file.c_str());
            if (param.rf) field.uRF.load(param.electric_field_rf_file.c_str());
            if (field.u.jmax != param.x_sampl || field.u.lmax != param.z_sampl)
                throw std::runtime_error("Pic: field file grid differs from x_sampl/z_sampl (regridding is not implemented)");
            field.upload();
        }
        else
        {
            field.boundary_solve_rf();     // the RF vacuum field is solved once (pic.cpp:180-187)
            if (!param.selfconsistent)
            {
                field.boundary_solve();
                field.reset();
            }
        }
        int64_t dummy;
        (void)dummy;
    }

    void check_params()
    {
        using std::cout;
        using std::endl;
        cout << endl << "**************** parameter validation ***************\n";
        cout << "V = " << param.V << " m3   dV = " << param.dV << " m3   dy = " << param.dy << " m" << endl << endl;
        for (auto J : speclist)
        {
            if (J->type == NEUTRAL) continue;
            cout << "**************** " << J->name << " parameters ***************\n";
            const double omega_p = sqrt(J->density * J->charge * J->charge / (J->mass * physconst::eps_0));
            const double lambda_D = sqrt(physconst::eps_0 * physconst::k_B * J->temperature / (J->density * J->charge * J->charge));
            const double v_thermal = sqrt(physconst::k_B * J->temperature / J->mass);
            const double maxdx = std::max(field.grid.dx, field.grid.dz), mindx = std::min(field.grid.dx, field.grid.dz);
            if (J->dt / J->lifetime > 0.2) cout << " *** WARNING dt > collisional  lifetime * 0.2 ***\n *** ";
            cout << " lifetime = " << J->lifetime << "    dt/lifetime = " << J->dt / J->lifetime << endl;
            if (J->dt * omega_p / (2 * M_PI) > 0.2) cout << " *** WARNING dt > plasma period * 0.2 ***\n *** ";
            cout << " omega_p  = " << omega_p << "    dt/period = " << omega_p / (2 * M_PI) * J->dt << endl;
            if (maxdx / lambda_D > 0.2) cout << " *** WARNING max grid spacing > lambda_D * 0.2 ***\n *** ";
            cout << " lambda_D = " << lambda_D << "    dx/lambda_D = " << maxdx / lambda_D << endl;
            if (J->density * param.dV < 20) cout << " *** WARNING particles per cell < 20 ***\n *** ";
            cout << " particles / cell = " << J->density * param.dV << endl;
            if (v_thermal * J->dt / mindx > 0.2) cout << " *** WARNING v_thermal*dt > min(dx) ***\n *** ";
            cout << " vth*dt / dx = " << v_thermal * J->dt / mindx << endl << endl;
        }
    }

    // the loader mini-language (pic.cpp:241-328)
    void run_initscript(const std::string& filename)
    {
        std::ifstream fr(filename.c_str());
        for (std::string line; std::getline(fr, line);)
        {
            if (!line.empty() && line[0] == '#') continue;
            std::istringstream words(line);
            std::vector<std::string> t;
            for (std::string w; words >> w;) t.push_back(w);
            if (t.empty()) continue;
            auto expect = [&](size_t n) {
                if (t.size() != n)
                    throw std::runtime_error("Pic::run_initscript: wrong number of parameters (" + std::to_string(t.size()) + ") to " + t[0] + "\n");
                if (!string2speciestype.count(t[1])) throw std::runtime_error("Pic::run_initscript: unrecognized species type \"" + t[1] + "\"\n");
                return speclist[(size_t)string2speciestype[t[1]]];
            };
            if (t[0] == "add_particles_bessel")
            {
                auto s = expect(6);
                s->add_particles_bessel(string2<int>(t[2]), string2<double>(t[3]), string2<double>(t[4]), string2<double>(t[5]));
                std::cout << s->name << " add bessel " << t[3] << " " << t[4] << " " << t[5] << "\n";
            }
            else if (t[0] == "add_particles_everywhere")
            {
                auto s = expect(3);
                s->add_particles_everywhere(string2<int>(t[2]));
                std::cout << s->name << " add everywhere " << t[2] << std::endl;
            }
            else if (t[0] == "add_particles_on_disk")
            {
                auto s = expect(6);
                s->add_particles_on_disk(string2<int>(t[2]), string2<double>(t[3]), string2<double>(t[4]), string2<double>(t[5]));
                std::cout << s->name << " add on disk " << t[2] << " " << t[3] << " " << t[4] << " " << t[5] << "\n";
            }
            else if (t[0] == "add_tracked_particle")
            {
                auto s = expect(7);
                s->add_tracked_particle(string2<double>(t[2]), string2<double>(t[3]), string2<double>(t[4]), string2<double>(t[5]), string2<double>(t[6]));
                has_tracked = true;
                std::cout << s->name << " add tracked particle " << t[2] << " " << t[3] << "  " << t[4] << " " << t[5] << " " << t[6] << std::endl;
            }
        }
        // the cell sort permutes slots, which would lose tracked particles: only sort when nothing is tracked
        gpu_check(mag2d_set_sort_interval(field.gpu, has_tracked ? 0 : -1));
    }

    void advance()
    {
        timer.start();
        gpu_check(mag2d_step(field.gpu, 1));
        gpu_check(mag2d_sync(field.gpu));
        timer.stop();
        iter++;
    }
    void advance_init() { gpu_check(mag2d_advance_init(field.gpu)); }

    void dist_reset()
    {
        nsampl = 0;
        timer.reset();
        for (auto s : speclist) s->dist_reset();
        field.u_reset();
    }
    void dist_sample()
    {
        nsampl++;
        for (auto s : speclist) s->dist_sample();
        field.u_sample();
    }
    void print_status(std::ostream& out = std::cout)
    {
        double probe_current_total = 0;
        for (auto s : speclist) probe_current_total += s->probe_current;
        field.download();
        out << iter << " " << timer.get_cpu_time() << " " << probe_current_total / nsampl * param.probe_length / param.dz << " "
            << field.grid.U_trap << " " << field.u[0][(int)(param.z_sampl * 4.0 / 7.5)] << std::endl;
        for (auto s : speclist) s->print_status();
    }
    void print_trace() { for (auto s : speclist) s->print_trace(); }
    void print_distribution() { for (auto s : speclist) s->print_distribution(); }
    void print_field() { field.u_print((param.output_dir + "/potential.dat").c_str()); }
    void save()
    {
        for (auto s : speclist)
            if (s->particle)                        // pic.cpp:462-470
            {
                s->save(param.output_dir + "/particles_" + s->name + ".dat");
                s->source5_save(param.output_dir + "/particles_source_" + s->name + ".dat");
            }
    }
    double step_seconds() { return timer.get_real_time(); }

  private:
    t_timer timer;
    bool has_tracked = false;
};
