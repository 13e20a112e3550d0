// species.hpp — t_particle, BaseSpecies, Species<D>: drop-in for the species surface of reference
// src/particles.hpp + src/particles.cpp.  The particles live on the device (SoA, cell-sorted); `particles`
// is a lazily synchronised AoS mirror with the reference's 64-byte record, kept for save()/load(), tracked
// particles and callers that edit the vector directly (src/test_MCC.cpp:62-68).
#pragma once
#include <cmath>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <random>
#include <string>
#include <vector>

#include "fields.hpp"
#include "parser.hpp"

class t_particle
{
  public:
    double x, y, z;
    double vx, vy, vz;
    double time_to_death;
    bool empty;
};
static_assert(sizeof(t_particle) == sizeof(mag2d_particle), "t_particle must keep the reference's 64-byte layout");

class BaseSpecies
{
  public:
    std::vector<t_particle> particles;      // host mirror; see download() / upload()
    SpeciesType type;
    std::string name;
    bool particle = true;
    double mass, charge, lifetime = INFINITY, temperature, polarizability, E_max, density, v_max, dt, t = 0;
    Field2D rho, rhoAverage;
    Histogram energy_dist, source_energy_dist, probe_energy_dist, probe_angular_dist, probe_angular_normalized_dist;
    double probe_current = 0, probe_charge = 0, probe_current_sum = 0;
    int nsampl = 0;
    std::vector<double> rates_by_species;
    int id = -1;               // index in the species list == device species index
    mag2d_ctx* gpu = nullptr;

    BaseSpecies(SpeciesParams* sp, Param& param, Fields& field, int index)
        : type(sp->type), name(sp->name), mass(sp->mass), charge(sp->charge), temperature(sp->temperature),
          polarizability(sp->polarizability), E_max(sp->E_max > 0. ? sp->E_max : sp->temperature * param.k_B / param.q_e * 10.0),
          density(sp->density), dt(sp->dt), rho(param.x_sampl, param.z_sampl, param.dx, param.dz),
          rhoAverage(param.x_sampl, param.z_sampl, param.dx, param.dz), energy_dist(200, 0.0, E_max), source_energy_dist(100, 0.0, E_max),
          probe_energy_dist(100, 0.0, E_max), probe_angular_dist(30, 0.0, M_PI * 0.5), probe_angular_normalized_dist(30, 0.0, M_PI * 0.5),
          id(index), gpu(field.gpu), p_param(&param)
    {
        v_max = sqrt(2.0 * physconst::k_B * temperature / mass);
        output.open((param.output_dir + "/" + name + ".dat").c_str());
    }
    virtual ~BaseSpecies() {}
    double EeV(double v) { return 0.5 * mass * v * v / p_param->q_e; }
    double veV(double e) { return sqrt(e * p_param->q_e / mass * 2.0); }

    // ---- particle store
    int n_particles()
    {
        int64_t alive = 0;
        gpu_check(mag2d_count(gpu, id, &alive, nullptr));
        return (int)alive;
    }
    // refresh the host mirror from the device (slot order, removed particles flagged empty)
    void download()
    {
        int64_t n = 0;
        gpu_check(mag2d_count(gpu, id, nullptr, &n));
        particles.resize((size_t)n);
        if (n) gpu_check(mag2d_particles_download(gpu, id, reinterpret_cast<mag2d_particle*>(particles.data()), n, &n));
    }
    // replace the device particles by the non-empty records of the host mirror
    void upload()
    {
        gpu_check(mag2d_particles_clear(gpu, id));
        if (!particles.empty()) gpu_check(mag2d_particles_upload(gpu, id, reinterpret_cast<const mag2d_particle*>(particles.data()), (int64_t)particles.size()));
    }
    void remove_all() { particles.clear(); gpu_check(mag2d_particles_clear(gpu, id)); }
    // binary checkpoint: int capacity; int n; n x t_particle (src/particles.cpp:32-59)
    void save(const std::string& filename)
    {
        download();
        std::ofstream fw(filename.c_str(), std::ios::out | std::ios::binary);
        if (!fw.is_open()) throw std::runtime_error("Species::save(): failed opening file");
        int cap = (int)particles.size(), n = 0;
        for (const t_particle& p : particles) n += p.empty ? 0 : 1;
        fw.write((char*)&cap, sizeof(int));
        fw.write((char*)&n, sizeof(int));
        for (const t_particle& p : particles)
            if (!p.empty) fw.write((const char*)&p, sizeof(t_particle));
    }
    void load(const std::string& filename)
    {
        std::ifstream fr(filename.c_str(), std::ios::in | std::ios::binary);
        if (!fr.is_open()) throw std::runtime_error("Species::load(): failed opening file");
        int cap = 0, n = 0;
        fr.read((char*)&cap, sizeof(int));
        fr.read((char*)&n, sizeof(int));
        particles.assign((size_t)n, t_particle());
        std::mt19937_64 gen(12345);
        std::uniform_real_distribution<double> uni(0.0, 1.0);
        for (int k = 0; k < n; k++)
        {
            fr.read((char*)&particles[k], sizeof(t_particle));
            if (!fr.good()) std::cerr << "Species::load(): read error\n";
            t_particle& p = particles[k];
            // particles that left the working area in an earlier non-selfconsistent run are put back (particles.cpp:77-83)
            if (p.x < p_param->x_min || p.x > p_param->x_max) p.x = (p_param->x_max - p_param->x_min) * uni(gen) + p_param->x_min;
            if (p.z < p_param->z_min || p.z > p_param->z_max) p.z = (p_param->z_max - p_param->z_min) * uni(gen) + p_param->z_min;
        }
        upload();
    }

    // the reservoir of the particle source (src/particles.cpp:109-139): int count + 64-byte t_particle records
    void source5_save(const std::string& filename)
    {
        std::ofstream fw(filename.c_str(), std::ios::out | std::ios::binary);
        if (!fw.is_open()) throw std::runtime_error("Species::source_save(): failed opening file");
        int64_t n64 = 0;
        gpu_check(mag2d_source_download(gpu, id, nullptr, 0, &n64));
        std::vector<mag2d_particle> buf((size_t)std::max<int64_t>(n64, 1));
        if (n64 > 0) gpu_check(mag2d_source_download(gpu, id, buf.data(), n64, &n64));
        int n = (int)n64;
        fw.write((char*)&n, sizeof(int));
        fw.write((const char*)buf.data(), sizeof(mag2d_particle) * (size_t)n);
    }
    void source5_load(const std::string& filename, unsigned int factor)
    {
        std::ifstream fr(filename.c_str(), std::ios::in | std::ios::binary);
        if (!fr.is_open())
        {
            std::cerr << " source5_load(): Warning: cannot open file " << filename << std::endl;
            return;
        }
        int n = 0;
        fr.read((char*)&n, sizeof(int));
        std::vector<mag2d_particle> buf((size_t)std::max(n, 1));
        fr.read((char*)buf.data(), sizeof(mag2d_particle) * (size_t)n);
        if (!fr.good()) std::cerr << "Species::load(): read error\n";
        gpu_check(mag2d_source_upload(gpu, id, factor, buf.data(), n));
    }

    // ---- diagnostics (src/particles.cpp:367-414)
    void energy_dist_compute()
    {
        std::vector<double> counts(energy_dist.N_hist());
        double stats[4];
        gpu_check(mag2d_energy_hist(gpu, id, energy_dist.N_hist(), energy_dist.Max(), counts.data(), stats));
        energy_dist.add_counts(counts.data(), stats);
    }
    // this species' charge grid in coulombs, from the device's fixed-point grid: rho = charge * W * 2^-32 (the reference
    // accumulates charge * weight per particle, Field2D.hpp:45-62; zero unless selfconsistent, particles.hpp:408)
    void rho_download()
    {
        std::vector<int64_t> fixed((size_t)rho.jmax * rho.lmax);
        gpu_check(mag2d_rho_fixed_download(gpu, id, fixed.data()));
        for (int i = 0; i < rho.jmax; i++)
            for (int j = 0; j < rho.lmax; j++) rho[i][j] = charge * ((double)fixed[(size_t)i * rho.lmax + j] * (1.0 / 4294967296.0));
    }
    void dist_sample()
    {
        energy_dist_compute();
        probe_current_sum += probe_current;
        rho_download();
        rhoAverage.add(rho);          // particles.cpp:386-393
        nsampl++;
    }
    void dist_reset()
    {
        energy_dist.reset(); source_energy_dist.reset(); probe_energy_dist.reset();
        probe_angular_dist.reset(); probe_angular_normalized_dist.reset(); rhoAverage.reset();
        nsampl = 0; probe_current_sum = 0; probe_current = 0;
    }
    void print_status(std::ostream& = std::cout)
    {
        double niter = 0;
        gpu_check(mag2d_species_get(gpu, id, 4, &niter));
        gpu_check(mag2d_species_get(gpu, id, 3, &t));
        output << (unsigned long)niter << " " << n_particles() << " " << energy_dist.mean_tot() << " " << t << std::endl;
    }
    void print_distribution()
    {
        const std::string base = p_param->output_dir + "/" + name;
        energy_dist.print((base + "_energy_dist.dat").c_str());
        probe_energy_dist.print((base + "_probe_energy_dist.dat").c_str());
        rhoAverage.print((base + "_rho.dat").c_str(), 1.0 / nsampl);
        probe_angular_dist.print((base + "_probe_angular_dist.dat").c_str());
        probe_angular_normalized_dist.print((base + "_probe_angular_normalized_dist.dat").c_str());
    }
    // tracked particles: slot indices whose trajectory goes to <name>_traj_<index>.dat (particles.cpp:5-17)
    void print_trace()
    {
        if (tracked.empty()) return;
        download();
        for (size_t k = 0; k < tracked.size(); k++)
        {
            const t_particle& p = particles[tracked[k]];
            *traj[k] << std::setprecision(10) << p.x << ' ' << p.z << ' ' << p.vx << ' ' << p.vz << ' ' << p.vy << std::endl;
        }
    }

  protected:
    Param* const p_param;
    std::ofstream output;
    std::vector<size_t> tracked;
    std::vector<std::ofstream*> traj;
};

template <int D>
class Species : public BaseSpecies
{
  public:
    Fields* field;
    Species(SpeciesParams* sp, Param& param, Fields& f, int index) : BaseSpecies(sp, param, f, index), field(&f) {}

    void advance() { gpu_check(mag2d_species_advance(gpu, id)); }
    void advance_init() { gpu_check(mag2d_species_advance_init(gpu, id)); }
    void accumulate() { gpu_check(mag2d_species_accumulate(gpu, id)); }
    // particle source (use_source; src/particles.cpp:1053-1080, 1158-1226): the reservoir lives on the device
    void source5_refresh(unsigned int factor)
    {
        gpu_check(mag2d_source_refresh(gpu, id, factor, p_param->V));
        int64_t n = 0;
        gpu_check(mag2d_source_download(gpu, id, nullptr, 0, &n));
        std::cout << name << " source initialized:  " << n << " particles  @  " << density << " particle/m3" << std::endl;
    }
    void source() { gpu_check(mag2d_species_source(gpu, id, nullptr)); }

    // loaders of the initscript language (src/particles.cpp:685-749); the uniform ones run on the device
    void add_particles_everywhere(int n) { gpu_check(mag2d_particles_generate(gpu, id, 0, n, 0, 0, 0, 0)); }
    void add_particles_on_disk(int n, double cx, double cy, double radius) { gpu_check(mag2d_particles_generate(gpu, id, 1, n, cx, cy, radius, 0)); }
    void add_monoenergetic_particles_on_cylinder_cylindrical(int n, double energy, double centerz, double radius, double height = 0.0)
    {
        gpu_check(mag2d_particles_generate(gpu, id, 2, n, energy, centerz, radius, height));
    }
    // density ~ J0(2.4048 r/R) on a disk: host rejection sampling, appended through the AoS upload
    void add_particles_bessel(int n, double cx, double cy, double radius)
    {
        static std::mt19937_64 gen(20241017);
        std::uniform_real_distribution<double> uni(0.0, 1.0);
        std::normal_distribution<double> nor(0.0, 1.0);
        const double root = 2.404825557695773;
        std::vector<mag2d_particle> batch;
        batch.reserve((size_t)n);
        for (int k = 0; k < n; k++)
        {
            double x, y, r;
            do
            {
                x = uni(gen) * 2 - 1.0;
                y = uni(gen) * 2 - 1.0;
                r = sqrt(x * x + y * y);
            } while (r > 1.0 || uni(gen) > std::cyl_bessel_j(0, r * root));
            x = x * radius + cx;
            y = y * radius + cy;
            if (x < p_param->x_min || x > p_param->x_max || y < p_param->z_min || y > p_param->z_max) continue;
            mag2d_particle p = {};
            p.x = x; p.z = y;
            p.vx = nor(gen) * v_max * M_SQRT1_2; p.vz = nor(gen) * v_max * M_SQRT1_2; p.vy = nor(gen) * v_max * M_SQRT1_2;
            batch.push_back(p);
        }
        if (!batch.empty()) gpu_check(mag2d_particles_upload(gpu, id, batch.data(), (int64_t)batch.size()));
    }
    void add_tracked_particle(double x, double y, double vx, double vy, double vz)
    {
        int64_t slots = 0;
        gpu_check(mag2d_count(gpu, id, nullptr, &slots));
        mag2d_particle p = {};
        p.x = x; p.z = y; p.y = 0;
        p.vx = vx; p.vy = vz; p.vz = vy;       // the reference stores the third argument pair swapped (particles.hpp:304-306)
        gpu_check(mag2d_particles_upload(gpu, id, &p, 1));
        tracked.push_back((size_t)slots);
        traj.push_back(new std::ofstream((p_param->output_dir + "/" + name + "_traj_" + std::to_string(slots) + ".dat").c_str()));
    }
};
