"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/mag2d_b200.h declares, validates configurations with the reference's messages, and refuses
to run without a CUDA device (no CPU fallback).  No compute calls are made here."""
import ctypes as C
import os
import re

import pytest

from mag2d_b200 import api
from mag2d_b200 import config as cfg
from mag2d_b200 import decks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from mag2d_b200.build import build
    return build()


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "mag2d_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(mag2d_[A-Za-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 40
    L = C.CDLL(built)
    missing = [name for name in declared if not hasattr(L, name)]
    assert not missing, missing
    assert L.mag2d_abi_version() == 4


def test_struct_layouts_match_header(built):
    assert api.PARTICLE_DTYPE.itemsize == 64          # reference t_particle (particles.hpp:26-33)
    assert C.sizeof(api.SpeciesDesc) == 56
    assert C.sizeof(api.InteractionDesc) == 48
    assert C.sizeof(api.GridDesc) == 3 * 4 + 3 * 4 + 9 * 8 + 4 * 4 + 8 + 3 * 8 + 2 * 4 + 3 * 8 + 2 * 8


def test_param_validation_uses_reference_messages(built, tmp_path):
    L = api.lib()
    d = decks.deck("c2", str(tmp_path), n_particles=10, x_sampl=21, z_sampl=21)
    p = cfg.read_config(d["config"])
    g = api.grid_desc_from_param(p)
    h = C.c_void_p()
    g.selfconsistent = 1     # rf = 1 in this deck
    assert L.mag2d_create(0, C.byref(g), None, C.byref(h)) != 0
    assert L.mag2d_last_error().decode() == "Param: selfconsistent rf trap not implemented\n"   # param.cpp:60-61
    g.selfconsistent, g.rf, g.electric_field_from_file = 1, 0, 1
    assert L.mag2d_create(0, C.byref(g), None, C.byref(h)) != 0
    assert "electric_field_from_file not implemented" in L.mag2d_last_error().decode()          # param.cpp:62-64
    g = api.grid_desc_from_param(p)
    g.coord, g.boundary = 1, 1
    assert L.mag2d_create(0, C.byref(g), None, C.byref(h)) != 0
    assert "only FREE boundary condition in cylindrical coords" in L.mag2d_last_error().decode()  # param.cpp:91-93
    with pytest.raises(cfg.ConfigError, match="selfconsistent rf trap"):
        cfg.read_config(d["config"], dict(selfconsistent=1))
    with pytest.raises(cfg.ConfigError, match="unrecognized mover"):
        cfg.read_config(d["config"], dict(mover="ADVANCE_LEAPFROG"))    # not selectable (param.cpp:98-103)
    with pytest.raises(cfg.ConfigError, match="MIRROR boundary"):
        cfg.read_config(d["config"], dict(boundary="MIRROR"))


def test_no_cpu_fallback(built, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    d = decks.deck("c1", str(tmp_path), n_particles=10)
    with pytest.raises(api.Mag2dError, match="no CUDA device"):
        api.Sim(d["config"], d["species_conf"])


def test_config_readers_shipped_quirks(tmp_path):
    # [section] blocks of old decks never shadow flat keys; integers go through the float parser
    f = tmp_path / "c.txt"
    f.write_text("coord = CARTESIAN # comment\nn_particles_total = 16e5\nrf = 0\n[ARGON]\npressure = ${* 133.0 1e-3}\nx_sampl = 7\n")
    p = cfg.read_config(str(f))
    assert p["n_particles_total"] == 16e5 and p["pressure"] == 133.0 and p["x_sampl"] == 100
    assert p["t_dist_sample"] == 1 and p["t_equilib"] == p["niter"] + 1
    s = tmp_path / "s.txt"
    s.write_text("SPECIES\n NAME A\n TYPE ION\n MASS 1e-27\n")
    sp, it = cfg.read_species(str(s))
    assert sp[0]["density"] == 0.0 and sp[0]["E_max"] == 0.0     # zero-initialised, unlike parser.cpp:45
    bad = tmp_path / "b.txt"
    bad.write_text("NAME A\n")
    with pytest.raises(cfg.ConfigError, match="unrecognized first config block"):
        cfg.read_species(str(bad))
