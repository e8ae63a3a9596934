"""Recipe for oracle/_ref/: the reference's own hot-path modules, taken unmodified from where they lie under
/root/reference, so that bench.py can time the REAL reference (not only the C port) on the GPU box's host cores.

    python oracle/build_ref.py            # needs /root/reference (this container); a no-op elsewhere

oracle/_ref/ is git-ignored (nothing of the reference enters the history) but is not gpurun-ignored, so it travels
to the GPU box like the built .so files.  The four files are pure Python on PyTorch (SURVEY.md finding 1) and
import only torch, math and each other: ha/ctc.py, ha/star.py, ha/transducer.py, ha/scan.py.
TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing under haloop_b200/ reads oracle/.
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/ha"
DST = os.path.join(HERE, "_ref", "ha")
FILES = ("ctc.py", "star.py", "transducer.py", "scan.py")


def build():
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    open(os.path.join(DST, "__init__.py"), "w").close()
    return True


def available():
    return all(os.path.exists(os.path.join(DST, f)) for f in FILES)


if __name__ == "__main__":
    print("oracle/_ref ready:", build())
