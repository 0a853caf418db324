"""Input batches for the shading known-answer tests (BSDF / light / sampler functions): random + edge cases.
Layouts: csrc/shading_kat.cuh (= oracle/ref_harness.cpp::ref_shading_batch)."""
import numpy as np

import scenes


def kat_scene(width=64, height=64):
    """small_mixed geometry with a material table that exercises every BSDF lobe: 1 diffuse, 2 diffuse + Phong,
    3 glass (delta, index 1.5), 4 mirror (delta, opaque), 5 Phong only, 6 black (continueProb 0), 7 dim diffuse just above
    the isBlack threshold, 8 diffuse + Phong + specular + dielectric (all four lobes)."""
    sc = scenes.small_mixed_scene(width, height)
    m = scenes.material
    mats = [m(), m(diffuse=(0.7, 0.6, 0.5)), m(diffuse=(0.3, 0.4, 0.2), phong=(0.4, 0.3, 0.3), phong_exp=12.0),
            m(specular=(1, 1, 1), index=1.5), m(specular=(0.9, 0.9, 0.9), index=-1.0), m(phong=(0.8, 0.8, 0.8), phong_exp=40.0),
            m(), m(diffuse=(0.004, 0.0035, 0.0011)), m(diffuse=(0.2, 0.2, 0.3), phong=(0.2, 0.1, 0.1), phong_exp=5.0, specular=(0.5, 0.5, 0.5), index=1.33)]
    sc.materials = np.asarray(mats, np.float32)
    sc.matid = np.where(sc.matid > 0, (np.arange(len(sc.matid)) % 8) + 1, sc.matid).astype(np.int32)
    return sc


def unit(v):
    v = np.asarray(v, np.float64)
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def bsdf_inputs(n, n_materials, seed, rand_last=False):
    """wi3 n3 matid wo3|rand3: random directions plus grazing wi / wo (|cos| around EPS), wo near the mirror direction,
    wo on the other side of the surface, axis-aligned normals."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nn = unit(rng.normal(size=(n, 3)))
    nn[::7] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, len(nn[::7]))] * np.where(rng.random(len(nn[::7])) < 0.5, -1, 1)[:, None].astype(np.float32)
    wi = unit(rng.normal(size=(n, 3)))
    wo = unit(rng.normal(size=(n, 3)))
    t = unit(np.cross(nn, rng.normal(size=(n, 3))))                         # tangent
    k = np.arange(n)
    g = (k % 5 == 1)                                                          # grazing wi: cos in [-3e-3, 3e-3]
    wi[g] = unit(t[g] + nn[g] * (rng.random((g.sum(), 1)) * 6e-3 - 3e-3))
    g = (k % 5 == 2)                                                          # grazing wo
    wo[g] = unit(t[g] + nn[g] * (rng.random((g.sum(), 1)) * 6e-3 - 3e-3))
    g = (k % 5 == 3)                                                          # wo close to the mirror direction of wi
    refl = 2 * np.sum(wi * nn, axis=1, keepdims=True) * nn - wi
    wo[g] = unit(refl[g] + rng.normal(size=(g.sum(), 3)) * 0.02)
    mat = rng.integers(1, n_materials, (n, 1)).astype(np.float32)
    last = rng.random((n, 3)).astype(np.float32) if rand_last else wo
    if rand_last:
        last[::11, 2] = np.float32(0.0); last[5::11, 2] = np.nextafter(np.float32(1.0), np.float32(0.0))      # lobe-selection edges
        last[3::13, 1] = np.float32(0.0)
    return np.concatenate([wi, nn, mat, last], axis=1).astype(np.float32)


def light_inputs(n, n_lights, seed, lo, hi):
    rng = np.random.Generator(np.random.PCG64(seed))
    lid = rng.integers(0, n_lights, (n, 1)).astype(np.float32)
    pos = (lo + (hi - lo) * (rng.random((n, 3)) * 1.4 - 0.2)).astype(np.float32)
    r = rng.random((n, 3)).astype(np.float32)
    r[::9, 0] = 0.0; r[4::9, 1] = 0.0
    return np.concatenate([lid, pos, r], axis=1).astype(np.float32)


def emit_inputs(n, n_lights, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    a = np.concatenate([rng.integers(0, n_lights, (n, 1)).astype(np.float32), rng.random((n, 6)).astype(np.float32)], axis=1)
    a[::10, 2] = 0.0                     # cos-hemisphere sample with z = sqrt(0): clamped to EPS (light.cpp:53)
    return a.astype(np.float32)


def radiance_inputs(n, n_lights, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.concatenate([rng.integers(0, n_lights, (n, 1)).astype(np.float32), unit(rng.normal(size=(n, 3)))], axis=1).astype(np.float32)


def fresnel_inputs(n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    c = (rng.random(n) * 2 - 1).astype(np.float32)
    c[::6] = (rng.random(len(c[::6])) * 4e-3 - 2e-3).astype(np.float32)
    idx = rng.choice(np.array([-1.0, 1.0, 1.33, 1.5, 2.4], np.float32), n)
    return np.stack([c, idx], axis=1).astype(np.float32)


def sampler_inputs(n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    r = rng.random((n, 3)).astype(np.float32)
    r[::8, 1] = 0.0; r[3::8, 0] = 0.0
    power = rng.choice(np.array([1.0, 5.0, 12.0, 40.0, 200.0], np.float32), (n, 1))
    tri = (rng.normal(size=(n, 9)) * 3).astype(np.float32)
    return np.concatenate([r, power, tri], axis=1).astype(np.float32)


def camera_inputs(n, width, height, spp, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.stack([rng.random(n), rng.random(n), rng.integers(0, height, n), rng.integers(0, width, n), rng.integers(0, spp, n)], axis=1).astype(np.float32)


def all_batches(sc, n, seed=1):
    lo = sc.data[sc.kind == 0].reshape(-1, 3).min(0); hi = sc.data[sc.kind == 0].reshape(-1, 3).max(0)
    nm, nl = len(sc.materials), len(sc.lights)
    return {
        0: bsdf_inputs(n, nm, seed), 1: bsdf_inputs(n, nm, seed + 1, rand_last=True), 2: bsdf_inputs(n, nm, seed + 2),
        3: light_inputs(n, nl, seed + 3, lo, hi), 4: emit_inputs(n, nl, seed + 4), 5: radiance_inputs(n, nl, seed + 5),
        6: fresnel_inputs(n, seed + 6), 7: sampler_inputs(n, seed + 7), 8: camera_inputs(n, sc.width, sc.height, 16, seed + 8),
    }


NAMES = {0: "BSDF::f", 1: "BSDF::sample", 2: "BSDF::pdf", 3: "AreaLight::illuminance", 4: "AreaLight::emit",
         5: "AreaLight::getRadiance", 6: "fresnelDielectric", 7: "samplers", 8: "camera sample"}


def compare(a, b, rtol, atol=1e-7):
    """Rows whose every entry agrees within rtol relative (+ a tiny absolute term for values around zero)."""
    return (np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)) + atol).all(axis=1)
