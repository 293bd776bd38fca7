"""examples/fanlin_stage.c: the boundary used from plain C, as a host binding would use it.  Without a GPU the program
must plan (host only) and then fail loudly with FANLIN_ENODEVICE; on a B200 its output is the oracle's within 1 LSB."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from synth import synth_image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory, fanlin):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    libdir = os.path.dirname(fanlin.lib_path())
    out = tmp_path_factory.mktemp("c_example") / "fanlin_stage"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "fanlin_stage.c"), "-L", libdir, "-lfanlin_device", f"-Wl,-rpath,{libdir}",
                    "-o", str(out)], check=True)
    return str(out)


def _write_ppm(path, img):
    h, w, c = img.shape
    with open(path, "wb") as f:
        f.write((b"P6" if c == 3 else b"P5") + b"\n# synthetic\n%d %d\n255\n" % (w, h) + np.ascontiguousarray(img).tobytes())


def _read_pam(path):
    raw = open(path, "rb").read()
    head, body = raw.split(b"ENDHDR\n", 1)
    f = dict(l.split(None, 1) for l in head.decode().splitlines()[1:] if l)
    w, h, c = int(f["WIDTH"]), int(f["HEIGHT"]), int(f["DEPTH"])
    return np.frombuffer(body, np.uint8).reshape(h, w, c)


def test_c_example_plans_on_the_host(exe):
    p = subprocess.run([exe, "--plan", "1920", "1080", "3", "w=300&h=200"], capture_output=True, text=True)
    assert p.returncode == 0 and "out 300x200 x 4 channels" in p.stdout and "resized 300x169" in p.stdout and "overlay at (0, 15)" in p.stdout
    assert "algorithmic bytes 6460800" in p.stdout
    p = subprocess.run([exe, "--plan", "3840", "2160", "4", "w=1618&h=1000&crop=true&blur=10"], capture_output=True, text=True)
    assert p.returncode == 0 and "resized 1778x1000, crop at (80, 0)" in p.stdout
    p = subprocess.run([exe, "--plan", "10", "10", "3", "w=+3"], capture_output=True, text=True)
    assert p.returncode == 1 and "failed to deserialize query string" in p.stderr  # FANLIN_EINVAL, the extractor's error


def test_c_example_fails_loudly_without_a_device(exe, tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _write_ppm(tmp_path / "in.ppm", synth_image(3, 40, 60, 3))
    p = subprocess.run([exe, str(tmp_path / "in.ppm"), "w=30&h=30", str(tmp_path / "out.pam")], capture_output=True, text=True)
    assert p.returncode == 5 and "no CPU fallback" in p.stderr  # FANLIN_ENODEVICE
    assert not os.path.exists(tmp_path / "out.pam")


@pytest.mark.gpu
@pytest.mark.parametrize("c,qs,kw", [
    (3, "w=300&h=200&rgb=32,32,32", dict(w=300, h=200, rgb=(32, 32, 32))),                      # the reference README's request
    (3, "w=120&h=80&crop=true&blur=10&grayscale=true", dict(w=120, h=80, crop=True, blur=10.0, grayscale=True)),
    (1, "w=64&h=64&inverse=true", dict(w=64, h=64, inverse=True)),
])
def test_c_example_matches_the_oracle(exe, tmp_path, c, qs, kw):
    img = synth_image(11 + c, 270, 480, c)
    _write_ppm(tmp_path / "in.pnm", img)
    p = subprocess.run([exe, str(tmp_path / "in.pnm"), qs, str(tmp_path / "out.pam")], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    got = _read_pam(tmp_path / "out.pam")
    want = O.process(img, **kw)
    assert got.shape == want.shape
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1
