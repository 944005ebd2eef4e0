"""The C ABI from plain C: tests/c_abi/consumer.c is compiled with gcc -std=c99 against include/gdl_b200.h (so the header is C-clean:
no C++ types, no torch) and linked with the host-compiled library of tests/hostemu, then run on malloc'ed buffers."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_plain_c_program_uses_the_abi(tmp_path):
    sys.path.insert(0, str(ROOT / "tests"))
    from hostemu import build
    lib = build.build()
    exe = tmp_path / "consumer"
    cmd = ["gcc", "-std=c99", "-D_POSIX_C_SOURCE=200112L", "-Wall", "-Werror", "-pedantic", f"-I{ROOT / 'include'}",
           str(ROOT / "tests" / "c_abi" / "consumer.c"), "-o", str(exe), f"-L{lib.parent}", f"-l:{lib.name}", f"-Wl,-rpath,{lib.parent}", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "consumer.c: ok" in r.stdout
