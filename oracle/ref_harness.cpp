// oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" window onto the UNMODIFIED reference (rouckas/mag2d) so that Python tests and
// bench.py's `--impl reference` / `cpu_baseline` legs can drive the reference's own CPU
// implementation of the hot path.  Nothing here is linked into, imported by, or executed from the
// product (mag2d_b200/); only tests/, __graft_entry__.smoke() and bench.py may load the resulting
// oracle/_ref/*.so.
//
// The reference sources are compiled where they lie under /root/reference/src by oracle/Makefile;
// this file only #includes the reference's pic.cpp (which, as in reference src/test.cpp:8, pulls in
// the whole class hierarchy) and forwards calls.  No reference source text is copied.
//
// Reference entry points driven from here:
//   Pic<D>::Pic / advance / advance_init / run_initscript   src/pic.cpp:127-189,330-384,241-328
//   Species<D>::advance / advance_position / advance_boundary src/particles.hpp:342-411
//   BaseSpecies::scatter / lifetime_init                       src/particles.cpp:208-365,142-170
//   Interaction::sigma_v                                      src/particles.hpp:61-70
//   Fields::E / boundary_solve / boundary_solve_rf / solve      src/fields.hpp:124-150, fields.cpp:278-353
//   t_random::uni / rnor / rexp / rot / deflect / radius      src/random.cpp:33-175
//
// The cylindrical driver needs `Species<CYLINDRICAL>::source()`, which the reference declares but
// never defines; an empty definition is supplied below (use_source is 0 in every
// shipped config so it is never reached with work to do).

// standard headers first, so the access-specifier override below never touches libstdc++
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include <tr1/cmath>
#include <sys/resource.h>
#include <sys/time.h>
#include <unistd.h>
#include <GetPot>

// the harness needs BaseSpecies::scatter (protected), ::niter (protected) and ::empty (private)
#define private public
#define protected public
#include "pic.cpp"
#undef private
#undef protected

template <>
void Species<CYLINDRICAL>::source() {}

namespace {

struct Handle
{
    GetPot config;
    Param* param = 0;
    Pic<CARTESIAN>* cart = 0;
    Pic<CYLINDRICAL>* cyl = 0;
    std::string err;
    t_random* rng() { return cart ? &cart->rnd : &cyl->rnd; }
    Fields* field() { return cart ? &cart->field : &cyl->field; }
    size_t nspecies() { return cart ? cart->speclist.size() : cyl->speclist.size(); }
    BaseSpecies* species(int i)
    {
        if (cart) return cart->speclist[(size_t)i];
        return cyl->speclist[(size_t)i];
    }
};

thread_local std::string g_err;

template <class F>
int guarded(Handle* h, F f)
{
    try { f(); return 0; }
    catch (std::exception& e) { g_err = e.what(); if (h) h->err = e.what(); return 1; }
    catch (...) { g_err = "unknown exception"; if (h) h->err = g_err; return 1; }
}

Field2D* pick_field(Handle* h, const char* which)
{
    std::string w(which);
    Fields* f = h->field();
    if (w == "u") return &f->u;
    if (w == "uRF") return &f->uRF;
    if (w == "rho") return &f->rho;
    if (w == "uAvg") return &f->uAvg;
    if (w.rfind("rho:", 0) == 0) return &h->species(atoi(w.c_str() + 4))->rho;
    return 0;
}

// the particle phase of Pic<D>::advance (pic.cpp:343-354) without the field solve (pic.cpp:333-341):
// per-species rho reset, Species::advance for every species, sum into field.rho
template <int D>
static void particle_phase(Pic<D>* pic)
{
    Param& param = pic->param;
    if (param.selfconsistent)
    {
        pic->field.reset();
        for (size_t i = 0; i < pic->speclist.size(); i++)
            if (pic->speclist[i]->particle && pic->speclist[i]->n_particles() > 0)
                pic->speclist[i]->rho.reset();
    }
    for (size_t i = 0; i < pic->speclist.size(); i++) pic->speclist[i]->advance();
    for (size_t i = 0; i < pic->speclist.size(); i++)
        if (pic->speclist[i]->particle && param.selfconsistent)
            pic->field.rho.add(pic->speclist[i]->rho);
}

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// overrides: "key=value;key=value" applied on top of the config file (e.g. "do_plot=0;mover=ADVANCE_BORIS")
void* ref_create(const char* config_file, const char* species_conf, const char* output_dir,
                 const char* overrides, unsigned seed)
{
    Handle* h = new Handle;
    int rc = guarded(h, [&] {
        h->config = GetPot(config_file);
        if (overrides)
        {
            std::string ov(overrides), item;
            std::istringstream is(ov);
            while (std::getline(is, item, ';'))
            {
                size_t eq = item.find('=');
                if (eq == std::string::npos) continue;
                h->config.set(item.substr(0, eq), item.substr(eq + 1));
            }
        }
        h->param = new Param(h->config);
        h->param->species_conf_file = species_conf;
        h->param->output_dir = output_dir;
        if (h->param->coord == CARTESIAN) h->cart = new Pic<CARTESIAN>(*h->param);
        else if (h->param->coord == CYLINDRICAL) h->cyl = new Pic<CYLINDRICAL>(*h->param);
        else throw std::runtime_error("ref_create: CARTESIAN3D has no runnable reference driver");
        h->rng()->initialize_seed(seed);
    });
    if (rc) { delete h; return 0; }
    return h;
}

void ref_destroy(void* hv)
{
    Handle* h = (Handle*)hv;
    if (!h) return;
    delete h->cart;
    delete h->cyl;
    delete h->param;
    delete h;
}

int ref_coord(void* hv) { return (int)((Handle*)hv)->param->coord; }

// Param dump in a fixed order (see tests/refharness.py PARAM_FIELDS)
int ref_param_get(void* hv, double* o)
{
    Param& p = *((Handle*)hv)->param;
    int k = 0;
    o[k++] = p.x_max; o[k++] = p.y_max; o[k++] = p.z_max;
    o[k++] = p.x_min; o[k++] = p.y_min; o[k++] = p.z_min;
    o[k++] = p.x_sampl; o[k++] = p.y_sampl; o[k++] = p.z_sampl;
    o[k++] = p.extern_field; o[k++] = p.electric_field_from_file; o[k++] = p.magnetic_field_const;
    o[k++] = p.Br; o[k++] = p.Bz; o[k++] = p.Bt;
    o[k++] = p.has_probe; o[k++] = p.probe_radius; o[k++] = p.probe_length; o[k++] = p.u_probe;
    o[k++] = p.n_particles_total; o[k++] = p.density_total;
    o[k++] = p.dx; o[k++] = p.dy; o[k++] = p.dz; o[k++] = p.V; o[k++] = p.dV;
    o[k++] = p.idx; o[k++] = p.idy; o[k++] = p.idz;
    o[k++] = p.pressure; o[k++] = p.neutral_temperature; o[k++] = p.macroparticle_factor;
    o[k++] = p.dt_elon; o[k++] = (double)p.niter;
    o[k++] = p.mover; o[k++] = p.coord; o[k++] = p.boundary; o[k++] = p.geometry;
    o[k++] = p.src_fact; o[k++] = p.selfconsistent; o[k++] = p.use_source; o[k++] = p.u_smooth;
    o[k++] = p.rf; o[k++] = p.rf_amplitude; o[k++] = p.rf_U0; o[k++] = p.rf_omega;
    o[k++] = p.particle_reload; o[k++] = (double)p.t_print; o[k++] = (double)p.t_print_dist;
    o[k++] = (double)p.t_dist_sample; o[k++] = (double)p.t_equilib; o[k++] = p.do_plot;
    o[k++] = p.neutral_density;
    return k;
}

int ref_n_species(void* hv) { return (int)((Handle*)hv)->nspecies(); }
const char* ref_species_name(void* hv, int i) { return ((Handle*)hv)->species(i)->name.c_str(); }

// [type, mass, charge, lifetime, temperature, E_max, density, v_max, dt, t, niter, n_particles, n_slots]
int ref_species_get(void* hv, int i, double* o)
{
    BaseSpecies* s = ((Handle*)hv)->species(i);
    int k = 0;
    o[k++] = s->type; o[k++] = s->mass; o[k++] = s->charge; o[k++] = s->lifetime;
    o[k++] = s->temperature; o[k++] = s->E_max; o[k++] = s->density; o[k++] = s->v_max;
    o[k++] = s->dt; o[k++] = s->t; o[k++] = (double)s->niter; o[k++] = s->n_particles();
    o[k++] = (double)s->particles.size();
    return k;
}
int ref_species_set(void* hv, int i, const char* what, double v)
{
    BaseSpecies* s = ((Handle*)hv)->species(i);
    std::string w(what);
    if (w == "niter") s->niter = (unsigned long)v;
    else if (w == "t") s->t = v;
    else if (w == "density") s->density = v;
    else if (w == "dt") s->dt = v;
    else if (w == "temperature") { s->temperature = v; s->v_max = sqrt(2.0 * physconst::k_B * v / s->mass); }
    else if (w == "lifetime") s->lifetime = v;
    else return 1;
    return 0;
}
int ref_lifetime_init(void* hv, int i)
{
    Handle* h = (Handle*)hv;
    return guarded(h, [&] { h->species(i)->lifetime_init(); });
}
int ref_species_rates(void* hv, int i, double* rates)
{
    BaseSpecies* s = ((Handle*)hv)->species(i);
    for (size_t k = 0; k < s->rates_by_species.size(); k++) rates[k] = s->rates_by_species[k];
    return (int)s->rates_by_species.size();
}
int ref_n_interactions(void* hv, int i, int target)
{
    return (int)((Handle*)hv)->species(i)->interactions_by_species[(size_t)target].size();
}
// [type, DE(J), rate, cutoff, n_table]
int ref_interaction_get(void* hv, int i, int target, int k, double* o)
{
    Interaction* I = ((Handle*)hv)->species(i)->interactions_by_species[(size_t)target][(size_t)k];
    o[0] = I->type; o[1] = I->DE; o[2] = I->rate; o[3] = I->cutoff;
    o[4] = 0;
    return 5;
}
double ref_sigma_v(void* hv, int i, int target, int k, double v)
{
    return ((Handle*)hv)->species(i)->interactions_by_species[(size_t)target][(size_t)k]->sigma_v(v);
}

int ref_run_initscript(void* hv, const char* path)
{
    Handle* h = (Handle*)hv;
    return guarded(h, [&] {
        if (h->cart) h->cart->run_initscript(path);
        else throw std::runtime_error("run_initscript: only Pic<CARTESIAN> instantiates the loaders");
    });
}

// Replace the particle array of species i by n particles; slot k holds input particle k.
// aos: n x 7 doubles (x,y,z,vx,vy,vz,time_to_death)
int ref_set_particles(void* hv, int i, int n, const double* aos)
{
    Handle* h = (Handle*)hv;
    return guarded(h, [&] {
        BaseSpecies* s = h->species(i);
        s->particles.clear();
        s->particles.resize((size_t)n);
        s->empty.clear();
        for (int k = 0; k < n; k++)
        {
            t_particle& p = s->particles[(size_t)k];
            const double* a = aos + 7 * (size_t)k;
            p.x = a[0]; p.y = a[1]; p.z = a[2]; p.vx = a[3]; p.vy = a[4]; p.vz = a[5];
            p.time_to_death = a[6];
            p.empty = false;
        }
    });
}
// out: n_slots x 8 doubles (x,y,z,vx,vy,vz,time_to_death,alive); returns n_slots written (<= max_slots)
int ref_get_particles(void* hv, int i, double* out, int max_slots)
{
    BaseSpecies* s = ((Handle*)hv)->species(i);
    int n = (int)std::min((size_t)max_slots, s->particles.size());
    for (int k = 0; k < n; k++)
    {
        const t_particle& p = s->particles[(size_t)k];
        double* a = out + 8 * (size_t)k;
        a[0] = p.x; a[1] = p.y; a[2] = p.z; a[3] = p.vx; a[4] = p.vy; a[5] = p.vz;
        a[6] = p.time_to_death; a[7] = p.empty ? 0.0 : 1.0;
    }
    return n;
}
int ref_n_slots(void* hv, int i) { return (int)((Handle*)hv)->species(i)->particles.size(); }

int ref_advance_init(void* hv)
{
    Handle* h = (Handle*)hv;
    return guarded(h, [&] { if (h->cart) h->cart->advance_init(); else h->cyl->advance_init(); });
}
int ref_advance(void* hv, int nsteps)
{
    Handle* h = (Handle*)hv;
    return guarded(h, [&] {
        for (int k = 0; k < nsteps; k++) { if (h->cart) h->cart->advance(); else h->cyl->advance(); }
    });
}
int ref_advance_particles(void* hv, int nsteps)
{
    Handle* h = (Handle*)hv;
    return guarded(h, [&] {
        for (int k = 0; k < nsteps; k++) { if (h->cart) particle_phase(h->cart); else particle_phase(h->cyl); }
    });
}
// push only (Species<D>::advance_position, no boundary / deposit / clock update)
int ref_advance_position(void* hv, int i, int init)
{
    Handle* h = (Handle*)hv;
    return guarded(h, [&] {
        if (h->cart)
        {
            Species<CARTESIAN>* s = h->cart->speclist[(size_t)i];
            if (init) s->advance_position_init(s->particles, false); else s->advance_position(s->particles, false);
        }
        else
        {
            Species<CYLINDRICAL>* s = h->cyl->speclist[(size_t)i];
            if (init) s->advance_position_init(s->particles, false); else s->advance_position(s->particles, false);
        }
    });
}
int ref_advance_boundary(void* hv, int i)
{
    Handle* h = (Handle*)hv;
    return guarded(h, [&] {
        if (h->cart) h->cart->speclist[(size_t)i]->advance_boundary();
        else h->cyl->speclist[(size_t)i]->advance_boundary();
    });
}
int ref_species_accumulate(void* hv, int i)
{
    Handle* h = (Handle*)hv;
    return guarded(h, [&] {
        if (h->cart) h->cart->speclist[(size_t)i]->accumulate();
        else h->cyl->speclist[(size_t)i]->accumulate();
    });
}
// ---- particle source (use_source), Cartesian only: Species<CARTESIAN>::source5_refresh / source (particles.cpp:1053-1080, 1158-1226)
int ref_source_refresh(void* hv, int i, unsigned factor)
{
    Handle* h = (Handle*)hv;
    return guarded(h, [&] {
        if (!h->cart) throw std::runtime_error("ref_source_refresh: Cartesian only");
        h->cart->speclist[(size_t)i]->source5_refresh(factor);
    });
}
int ref_source(void* hv, int i)
{
    Handle* h = (Handle*)hv;
    return guarded(h, [&] {
        if (!h->cart) throw std::runtime_error("ref_source: Cartesian only");
        h->cart->speclist[(size_t)i]->source();
    });
}
int ref_source_n(void* hv, int i) { return (int)((Handle*)hv)->species(i)->source2_particles.size(); }
int ref_source_get(void* hv, int i, double* out, int max_slots)
{
    BaseSpecies* s = ((Handle*)hv)->species(i);
    int n = (int)std::min((size_t)max_slots, s->source2_particles.size());
    for (int k = 0; k < n; k++)
    {
        const t_particle& p = s->source2_particles[(size_t)k];
        double* a = out + 8 * (size_t)k;
        a[0] = p.x; a[1] = p.y; a[2] = p.z; a[3] = p.vx; a[4] = p.vy; a[5] = p.vz;
        a[6] = p.time_to_death; a[7] = p.empty ? 0.0 : 1.0;
    }
    return n;
}
void ref_srand(unsigned seed) { srand(seed); }     // source() draws its lateral shifts from libc rand() (particles.cpp:1177)
// seconds for nsteps of the chosen phase, measured with the reference's own t_timer (timer.hpp:15-27):
// returns wall-clock seconds, *cpu_seconds gets the getrusage user time test_MCC.cpp:86 reports
double ref_time_advance(void* hv, int nsteps, int particles_only, double* cpu_seconds)
{
    Handle* h = (Handle*)hv;
    t_timer timer;
    timer.reset();
    int rc = guarded(h, [&] {
        for (int k = 0; k < nsteps; k++)
        {
            timer.start();
            if (particles_only) { if (h->cart) particle_phase(h->cart); else particle_phase(h->cyl); }
            else { if (h->cart) h->cart->advance(); else h->cyl->advance(); }
            timer.stop();
        }
    });
    if (cpu_seconds) *cpu_seconds = timer.get_cpu_time();
    return rc ? -1.0 : timer.get_real_time();
}

int ref_grid_dims(void* hv, int* M, int* N)
{
    Fields* f = ((Handle*)hv)->field();
    *M = f->grid.M; *N = f->grid.N;
    return 0;
}
int ref_get_field(void* hv, const char* which, double* out)
{
    Handle* h = (Handle*)hv;
    Fields* f = h->field();
    const int M = f->grid.M, N = f->grid.N;
    std::string w(which);
    if (w == "mask") { for (int i = 0; i < M; i++) for (int j = 0; j < N; j++) out[i * N + j] = f->grid.mask[i][j]; return 0; }
    if (w == "voltage") { for (int i = 0; i < M; i++) for (int j = 0; j < N; j++) out[i * N + j] = f->grid.voltage[i][j]; return 0; }
    Field2D* F = pick_field(h, which);
    if (!F) { g_err = "ref_get_field: unknown field"; return 1; }
    for (int i = 0; i < F->jmax; i++) for (int j = 0; j < F->lmax; j++) out[i * F->lmax + j] = (*F)[i][j];
    return 0;
}
int ref_set_field(void* hv, const char* which, const double* in)
{
    Handle* h = (Handle*)hv;
    Field2D* F = pick_field(h, which);
    if (!F) { g_err = "ref_set_field: unknown field"; return 1; }
    for (int i = 0; i < F->jmax; i++) for (int j = 0; j < F->lmax; j++) (*F)[i][j] = in[i * F->lmax + j];
    return 0;
}
int ref_field_op(void* hv, const char* op)
{
    Handle* h = (Handle*)hv;
    std::string w(op);
    return guarded(h, [&] {
        Fields* f = h->field();
        if (w == "boundary_solve") f->boundary_solve();
        else if (w == "boundary_solve_rf") f->boundary_solve_rf();
        else if (w == "solve") f->solve();
        else if (w == "reset") f->reset();
        else if (w == "u_smooth") f->u_smooth();
        else throw std::runtime_error("ref_field_op: unknown op " + w);
    });
}
// BaseSpecies::save / load (src/particles.cpp:32-93): the binary checkpoint either side of the path
int ref_species_save(void* hv, int i, const char* path)
{
    Handle* h = (Handle*)hv;
    return guarded(h, [&] { h->species(i)->save(path); });
}
int ref_species_load(void* hv, int i, const char* path)
{
    Handle* h = (Handle*)hv;
    return guarded(h, [&] { h->species(i)->load(path); });
}
// Field2D::load (src/Field2D.cpp:46-130) into a scratch Field2D: dimensions, origin, spacing and values of what was read
int ref_field2d_load(void* hv, const char* path, double* info6, double* values, int max_values)
{
    Handle* h = (Handle*)hv;
    return guarded(h, [&] {
        Field2D F(2, 2, 1.0, 1.0);
        F.load(path);
        info6[0] = F.jmax; info6[1] = F.lmax; info6[2] = F.GetXMin(); info6[3] = F.GetYMin(); info6[4] = F.GetXMax(); info6[5] = F.GetYMax();
        if (F.jmax * F.lmax > max_values) throw std::runtime_error("ref_field2d_load: buffer too small");
        for (int a = 0; a < F.jmax; a++) for (int b = 0; b < F.lmax; b++) values[a * F.lmax + b] = F[a][b];
    });
}
// BaseSpecies::energy_dist_compute (src/particles.cpp:408-414) into a fresh Histogram of the species' own shape
// (particles.hpp:169: 200 bins over [0, E_max)); out = n_hist bins, stats = n_val, mean, mean_tot, norm
int ref_energy_hist(void* hv, int i, int n_hist, double* out, double* stats)
{
    Handle* h = (Handle*)hv;
    return guarded(h, [&] {
        BaseSpecies* S = h->species(i);
        S->energy_dist.reset();
        S->energy_dist_compute();
        if (S->energy_dist.N_hist() != n_hist) throw std::runtime_error("ref_energy_hist: bin count differs");
        for (int k = 0; k < n_hist; k++) out[k] = S->energy_dist[k];
        stats[0] = S->energy_dist.N_val();
        stats[1] = S->energy_dist.mean();
        stats[2] = S->energy_dist.mean_tot();
        stats[3] = S->energy_dist.norm();
        stats[4] = S->energy_dist.Min();
        stats[5] = S->energy_dist.Max();
    });
}
int ref_field_E(void* hv, int n, const double* x, const double* z, double time, double* Ex, double* Ez)
{
    Handle* h = (Handle*)hv;
    return guarded(h, [&] { for (int k = 0; k < n; k++) h->field()->E(x[k], z[k], Ex[k], Ez[k], time); });
}
/* Fields::B (fields.hpp:152-177): the constants or the table read by Fields::load_magnetic_field in the Pic constructor */
int ref_field_B(void* hv, int n, const double* x, const double* z, double* Br, double* Bz, double* Bt)
{
    Handle* h = (Handle*)hv;
    return guarded(h, [&] { for (int k = 0; k < n; k++) h->field()->B(x[k], z[k], Br[k], Bz[k], Bt[k]); });
}
/* the loaded tables themselves: dims (rsampl, zsampl, dx, dz, rmin, zmin), then the data with which = "Br" / "Bz" */
int ref_btable_info(void* hv, double* o)
{
    Handle* h = (Handle*)hv;
    return guarded(h, [&] {
        Field2D& f = h->field()->Br;
        o[0] = f.jmax; o[1] = f.lmax; o[2] = f.GetDx(); o[3] = f.GetDy(); o[4] = f.GetXMin(); o[5] = f.GetYMin();
    });
}
int ref_btable_get(void* hv, int which, double* out)
{
    Handle* h = (Handle*)hv;
    return guarded(h, [&] {
        Field2D& f = which == 0 ? h->field()->Br : h->field()->Bz;
        for (int i = 0; i < f.jmax; i++)
            for (int j = 0; j < f.lmax; j++) out[(size_t)i * f.lmax + j] = f[i][j];
    });
}
int ref_field_accumulate(void* hv, const char* which, double charge, int n, const double* x, const double* z)
{
    Handle* h = (Handle*)hv;
    return guarded(h, [&] {
        Field2D* F = pick_field(h, which);
        if (!F) throw std::runtime_error("ref_field_accumulate: unknown field");
        for (int k = 0; k < n; k++) F->accumulate(charge, x[k], z[k]);
    });
}
int ref_is_free(void* hv, int n, const double* x, const double* z, int* out)
{
    Handle* h = (Handle*)hv;
    for (int k = 0; k < n; k++) out[k] = h->field()->grid.is_free(x[k], z[k]) ? 1 : 0;
    return 0;
}

// scatter n velocity triples (vx, vz, vy as stored in t_particle: v[3k]=vx, v[3k+1]=vy, v[3k+2]=vz) in place
int ref_scatter(void* hv, int i, int n, double* v)
{
    Handle* h = (Handle*)hv;
    return guarded(h, [&] {
        BaseSpecies* s = h->species(i);
        t_particle p;
        p.x = p.y = p.z = 0; p.time_to_death = 0; p.empty = false;
        for (int k = 0; k < n; k++)
        {
            p.vx = v[3 * k]; p.vy = v[3 * k + 1]; p.vz = v[3 * k + 2];
            s->scatter(p);
            v[3 * k] = p.vx; v[3 * k + 1] = p.vy; v[3 * k + 2] = p.vz;
        }
    });
}

void ref_rng_seed(void* hv, unsigned seed) { ((Handle*)hv)->rng()->initialize_seed(seed); }
void ref_rng_draw(void* hv, const char* what, int n, double* out)
{
    t_random* r = ((Handle*)hv)->rng();
    std::string w(what);
    for (int k = 0; k < n; k++)
    {
        if (w == "uni") out[k] = r->uni();
        else if (w == "rnor") out[k] = r->rnor();
        else if (w == "rexp") out[k] = r->rexp();
        else if (w == "iuni") out[k] = r->iuni();
        else if (w == "radius") out[k] = r->radius();
    }
}
// rot(len): out 3 doubles per draw; deflect(angle, v): in/out 3 doubles
void ref_rng_rot(void* hv, double len, int n, double* out)
{
    t_random* r = ((Handle*)hv)->rng();
    for (int k = 0; k < n; k++) r->rot(len, out[3 * k], out[3 * k + 1], out[3 * k + 2]);
}
void ref_rng_deflect(void* hv, double angle, int n, double* v)
{
    t_random* r = ((Handle*)hv)->rng();
    for (int k = 0; k < n; k++) r->deflect(angle, v[3 * k], v[3 * k + 1], v[3 * k + 2]);
}
}  // extern "C"
