#!/usr/bin/env python3
"""Collect the reference's own regression fixtures for the hot path into tests/golden/ (run in the build container,
where /root/reference exists; the GPU box only sees the committed copies).

  example/limb : limb.ctl, obs.tab (66 rays), atm.tab (91 levels, 5 gases), rad.org (golden output of formod)
  example/nadir: nadir.ctl, obs.tab (only the first 90 rays = time 0), atm.tab, rad.org

The emissivity tables that produced the radiance/transmittance columns of rad.org are missing from the reference
checkout (.MISSING_LARGE_BLOBS), so only the geometry columns 1-8 and 10 of rad.org are usable as known answers
(they pin the ray tracer and tangent_point).  Filter files are not copied: tests use synthetic boxcar filters.
"""
import os
import shutil

REF = os.environ.get("JURASSIC_REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def data_rows(path):
    return [l for l in open(path) if l.strip() and not l.startswith("#")]


def main():
    for case, ctl in (("limb", "limb.ctl"), ("nadir", "nadir.ctl")):
        src = os.path.join(REF, "example", case)
        dst = os.path.join(HERE, case)
        os.makedirs(dst, exist_ok=True)
        shutil.copyfile(os.path.join(src, ctl), os.path.join(dst, ctl))
        shutil.copyfile(os.path.join(src, "atm.tab"), os.path.join(dst, "atm.tab"))
        shutil.copyfile(os.path.join(src, "rad.org"), os.path.join(dst, "rad.org"))
        if case == "limb":
            shutil.copyfile(os.path.join(src, "obs.tab"), os.path.join(dst, "obs.tab"))
        else:
            rows = data_rows(os.path.join(src, "obs.tab"))[:90]
            with open(os.path.join(dst, "obs.tab"), "w") as f:
                f.write("# first 90 rays (time 0) of the reference's example/nadir/obs.tab\n")
                f.writelines(rows)
        for fn in os.listdir(dst):
            os.chmod(os.path.join(dst, fn), 0o644)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
