#!/usr/bin/env python3
"""Regenerate the committed golden fixtures from the REFERENCE BINARY (authoring container only).

  kat_<group>.bin.gz : known-answer vectors harvested from /root/reference/centos_x64/appencoder's own
                       scalar kernels by oracle/kat/harvest.c (LD_PRELOAD, absolute addresses)
Needs /root/reference; run `make -C oracle` first.  Deterministic (fixed xorshift seed in harvest.c).
"""
import gzip, os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
root = os.path.dirname(os.path.dirname(here))
ref = os.path.join(root, "oracle", "_ref")
subprocess.check_call(["make", "-s", "-C", os.path.join(root, "oracle")])
for g in ("sad", "transform", "interp", "loop", "tables"):
    tmp = "/tmp/kat_%s.bin" % g
    env = dict(os.environ, KS_KAT_OUT=tmp, KS_KAT_WHAT=g, LD_PRELOAD=os.path.join(ref, "kat_harvest.so"))
    subprocess.check_call([os.path.join(ref, "appencoder"), "-v"], env=env, stdout=subprocess.DEVNULL)
    with open(tmp, "rb") as f, gzip.GzipFile(os.path.join(here, "kat_%s.bin.gz" % g), "wb", mtime=0) as z:
        z.write(f.read())
    print(g, os.path.getsize(os.path.join(here, "kat_%s.bin.gz" % g)))
