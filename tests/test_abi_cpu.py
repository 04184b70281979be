"""CPU: the C-ABI library builds, loads and exports every symbol ``include/ffb200.h`` declares; ctypes structs
match the C layouts; the product refuses to run without a GPU instead of falling back."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ffb200.h")


@pytest.fixture(scope="module")
def nat():
    from fireflies_b200 import _build, _native
    _build.build()
    return _native


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"FFB_API\s+[\w\s\*]+?\b(ffb_\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(nat):
    names = declared_symbols()
    assert len(names) >= 20
    lib = nat.lib()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in ffb200.h but not exported"
        assert n in nat.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(nat.SIGNATURES) == names
    assert lib.ffb_version() == 100


def test_struct_layouts_match_header(nat, tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ffb200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                    "sizeof(ffb_splat_desc),sizeof(ffb_sampler),sizeof(ffb_entity),sizeof(ffb_mesh_table),sizeof(ffb_post_desc),"
                    "offsetof(ffb_sampler,vmin),offsetof(ffb_entity,world),offsetof(ffb_mesh_table,frames),offsetof(ffb_post_desc,seed));return 0;}")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(nat.SplatDesc), C.sizeof(nat.Sampler), C.sizeof(nat.Entity), C.sizeof(nat.MeshTable), C.sizeof(nat.PostDesc),
            nat.Sampler.vmin.offset, nat.Entity.world.offset, nat.MeshTable.frames.offset, nat.PostDesc.seed.offset]
    assert got == want


def test_argument_errors_do_not_need_a_gpu(nat):
    lib = nat.lib()
    d = nat.SplatDesc(0, 4, 8, 8, 1.0, 4, 5, 0)
    assert lib.ffb_splat_workspace_bytes(C.byref(d)) == 0
    assert b"positive" in lib.ffb_last_error_string()
    d = nat.SplatDesc(2, 4096, 2048, 2048, 100.0, 4, 5, 8192)
    assert lib.ffb_splat_workspace_bytes(C.byref(d)) > 2 * 4096 * 32
    assert lib.ffb_reduce_over_samples(None, 1, 1, None, None) == -1


def test_no_cpu_fallback():
    import fireflies_b200 as ff
    R = ff.graphics.rasterization
    with pytest.raises(RuntimeError, match="no CPU path"):
        R.baked_sum(torch.rand(4, 2), torch.tensor([9.0]), torch.tensor([32, 32]))
    with pytest.raises(RuntimeError, match="no CPU path"):
        R.rasterize_points(torch.rand(4, 2).cuda() if torch.cuda.is_available() else torch.rand(4, 2), 4.0, [8, 8], device="cpu")
    with pytest.raises(RuntimeError):
        ff.utils.math.transform_points(torch.rand(5, 3), torch.eye(4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fireflies_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


def test_api_surface_matches_reference_names():
    import fireflies_b200 as ff
    for mod, names in {
        ff.graphics.rasterization: ["rasterize_points", "softor", "sum", "baked_sum", "baked_sum_2", "baked_softor",
                                    "baked_softor_2", "rasterize_points_baked_softor", "rasterize_points_baked_sum",
                                    "rasterize_points_in_non_ndc", "rasterize_lines", "rasterize_depth", "subsampled_point_raster"],
        ff.utils.math: ["getYawTransform", "getPitchTransform", "getRollTransform", "randomBetweenTensors", "toMat4x4",
                        "transform_points", "transform_directions", "convert_points_to_homogeneous"],
        ff.sampling: ["Sampler", "UniformSampler", "UniformScalarToVec3Sampler", "GaussianSampler", "AnimationSampler",
                      "UniformIntegerSampler", "NoiseTextureLerpSampler"],
        ff.entity: ["Transformable", "Mesh", "Curve"],
        ff.projection: ["Camera", "Laser"],
        ff.postprocessing: ["BasePostProcessingFunction", "PostProcessor", "GaussianBlur", "WhiteNoise", "ApplySilhouette"],
    }.items():
        for n in names:
            assert hasattr(mod, n), (mod.__name__, n)
    for n in ["rotate_x", "rotate_y", "rotate_z", "rotate", "translate_x", "translate", "set_world", "world", "setParent",
              "set_centroid", "randomize", "train", "eval", "add_float_key", "add_vec3_key", "add_vec3_sampler",
              "nonRandomizedWorld", "get_randomized_float_attributes"]:
        assert hasattr(ff.entity.Transformable, n), n
    for n in ["scale_x", "scale", "get_randomized_vertices", "add_animation_func", "add_train_animation_from_obj",
              "sample_animation", "get_vertices", "set_vertices"]:
        assert hasattr(ff.entity.Mesh, n), n
    for n in ["generate_uniform_rays", "projectRaysToNDC", "projectNDCPointsToWorld", "generateTexture", "clamp_to_fov",
              "normalize_rays", "save", "rays", "originPerRay", "generate_uniform_rays_by_count", "generate_random_rays",
              "initRandomRays", "randomize_laser_out_of_bounds", "randomize_camera_out_of_bounds", "render_epipolar_lines"]:
        assert hasattr(ff.projection.Laser, n), n
    assert ff.scene is ff.Scene
