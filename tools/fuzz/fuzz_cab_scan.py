"""Random mutations of the header area of the fixture cabinets: msgpu_cab_scan() against the reference's mscab_decompressor::open()
(oracle/_ref/ref_cabx = cabd.c + system.c) - error code and file table.  usage: fuzz_cab_scan.py [seed] [cases]
Development aid (CPU only): TEST INFRASTRUCTURE, like everything that loads oracle/.  Run from the repository root."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import json, subprocess, tempfile, shutil
import numpy as np
from libmspack_b200 import cab
CABDIR = os.path.join(ROOT, 'tests', 'golden', 'cab'); CABX = os.path.join(ROOT, 'oracle', '_ref', 'ref_cabx')      # make -C oracle cabx
names=[e["name"] for e in json.load(open(CABDIR+'/manifest.json')) if os.path.getsize(CABDIR+'/'+e["name"]) < 200000]
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
tmp=tempfile.mkdtemp(); bad=0; n=0
for it in range(int(sys.argv[2]) if len(sys.argv)>2 else 300):
    name=names[int(rng.integers(0,len(names)))]
    img=bytearray(open(CABDIR+'/'+name,'rb').read())
    hdr_end=min(len(img), 0x24+400)
    for _ in range(int(rng.integers(1,4))):
        p=int(rng.integers(0,hdr_end)); img[p]=int(rng.integers(0,256)) if rng.random()<0.5 else img[p]^(1<<int(rng.integers(0,8)))
    if rng.random()<0.15: img=img[:int(rng.integers(1,len(img)))]
    path=os.path.join(tmp,'x.cab'); open(path,'wb').write(bytes(img))
    out=os.path.join(tmp,'o'); os.makedirs(out,exist_ok=True)
    try:
        lines=subprocess.run([CABX,path,out],stdout=subprocess.PIPE,stderr=subprocess.DEVNULL,timeout=60).stdout.decode('ascii','replace').split('\n')
    except subprocess.TimeoutExpired:
        continue
    if not lines or not lines[0].startswith('open'): continue
    ref_open=int(lines[0].split()[1]); ref_files=[tuple(int(x) for x in l.split()) for l in lines[1:] if l.strip()]
    n+=1
    try:
        plan=cab.scan(bytes(img)); mine_open=0
    except cab.CabError as e:
        mine_open=e.code; plan=None
    ok = (mine_open==ref_open)
    if ok and plan is not None:
        if len(plan.files)!=len(ref_files): ok=False
        else:
            for rec,f in zip(ref_files,plan.files):
                if (int(f["offset"]),int(f["length"]))!=(rec[2],rec[3]): ok=False
                if rec[1]>=0 and int(f["folder"])!=0xFFFFFFFF and int(f["folder"])!=rec[1]: ok=False
    if not ok:
        bad+=1
        if bad<=8:
            print("MISMATCH", name, "ref open", ref_open, "mine", mine_open, "files", len(ref_files), None if plan is None else len(plan.files))
            open(f'/tmp/cabfuzz_bad_{bad}.cab','wb').write(bytes(img))
shutil.rmtree(tmp)
print("cab scan fuzz:", n, "cases,", bad, "mismatches")
