"""The command line (csrc/cli_main.cpp), which keeps the reference's surface `<binary> scenes/x.toml` (main.rs:43-51) and
its progress lines.  What can be checked without a GPU: argument handling, the reference's messages for a missing
argument / file, the host-side lines printed before the device is touched, and the loud refusal without a device."""
import os
import subprocess

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def cli(lr, assets):
    from lumillyrender_b200 import build
    path = build.build_cli()
    assert path and os.path.exists(path)
    return path


def run(cli, *args, cwd=None):
    return subprocess.run([cli, *args], capture_output=True, text=True, cwd=cwd or ROOT, timeout=120)


def test_missing_argument_is_the_reference_message(cli):
    r = run(cli)
    assert r.returncode == 2 and "Path for .toml must be specified." in r.stderr      # main.rs:47-49
    assert r.stdout.startswith("start: ")                                              # main.rs:44


def test_bad_flags_exit_2(cli):
    assert run(cli, "--spp").returncode == 2
    assert run(cli, "scenes/primitive.toml", "--resolution", "banana").returncode == 2
    assert run(cli, "scenes/primitive.toml", "extra.toml").returncode == 2


def test_missing_scene_file_is_an_error_not_a_crash(cli):
    r = run(cli, "scenes/does-not-exist.toml")
    assert r.returncode == 1 and "is not found" in r.stderr                           # description.rs:34
    assert "loading: scenes/does-not-exist.toml" in r.stdout


def test_host_side_lines_then_loud_refusal_without_a_gpu(cli):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = run(cli, "scenes/primitive.toml", "--spp", "2", "--resolution", "32x32")
    assert r.returncode == 1
    assert "resolution: 32x32" in r.stdout and "spp: 2" in r.stdout                   # main.rs:52-59, printed before the device is touched
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cli_renders_and_writes_the_image(cli, tmp_path):
    r = run(cli, os.path.join(ROOT, "scenes", "primitive.toml"), "--spp", "4", "--resolution", "64x64", "--assets", ROOT, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    for line in ("start: ", "loading: ", "resolution: 64x64", "spp: 4", "integrator: pt", "polygons: ", "bvh construction: ", "saving...", "end: ", "elapse: "):
        assert line in r.stdout, line
    files = os.listdir(os.path.join(str(tmp_path), "images"))
    assert len(files) == 1 and files[0].startswith("image_") and files[0].endswith("_4.png")   # main.rs:147-169
