"""container_fuzz.py SEED COUNT WORKDIR -- child process of tests/test_container_fuzz_cpu.py.

Writes a small .mcraw (four frames of both formats, audio with and without timestamps), then COUNT mutated copies
(random bytes, extreme 32/64-bit values in item headers and indexes, truncations) and walks each through this repo's
drop-in motioncam::Decoder on the CPU: open, frame list, metadata, audio (both loaders), loadFrame of the first and
last frame (which ends in an IOException without a GPU, after the read + JSON parse that is the point here).
Every failure must surface as an exception; the process prints "ok {...}" at the end.  It runs as a child so that an
abort or a segmentation fault is a test failure rather than the end of pytest; the address space is capped at 8 GiB
so that a size field taken at face value shows up as a failed allocation."""
import os, sys, struct, resource
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from motioncam_decoder_b200 import hostapi, testvec as tv
if not os.environ.get('MCRAW_FUZZ_NO_RLIMIT'):      # (an AddressSanitizer build needs its shadow mappings)
    resource.setrlimit(resource.RLIMIT_AS, (8<<30, 8<<30))
seed=int(sys.argv[1]); N=int(sys.argv[2])
rng=np.random.default_rng(seed)
d=sys.argv[3]
frames=[]
for k in range(4):
    w,h=64+32*k,8
    img=tv.gen_photon(w,h,1023,seed=10+k)
    legacy=bool(k&1)
    frames.append({"timestamp":1000+10*k,"data":tv.encode_legacy(img) if legacy else tv.encode_current(img),"width":w,"height":h,"compressionType":6 if legacy else 7})
audio=[(123456789, rng.integers(-32768,32767,960,dtype=np.int16)),(None, rng.integers(-32768,32767,481,dtype=np.int16)),(223456789, rng.integers(-32768,32767,2,dtype=np.int16))]
base=os.path.join(d,f'base{seed}.mcraw'); tv.write_mcraw(base,frames,audio)
raw=bytearray(open(base,'rb').read())
L=len(raw)
special=[0,1,0xFF,0x7F,0x80,0xFE]
stats={'open_fail':0,'open_ok':0,'frame_err':0,'audio_err':0}
for it in range(N):
    b=bytearray(raw)
    kind=rng.integers(0,6)
    if kind==0:   # random bytes anywhere
        for _ in range(int(rng.integers(1,8))):
            b[int(rng.integers(0,L))]=int(rng.integers(0,256))
    elif kind==1: # tail region (indexes)
        for _ in range(int(rng.integers(1,6))):
            b[L-1-int(rng.integers(0,min(L,260)))]=int(rng.choice(special))
    elif kind==2: # truncate
        b=b[:int(rng.integers(0,L))]
    elif kind==3: # 32-bit word to an extreme value at random aligned-ish position
        pos=int(rng.integers(0,L-4)); b[pos:pos+4]=struct.pack('<I',int(rng.choice([0xFFFFFFFF,0x7FFFFFFF,0x80000000,0,1,L,L*2])))
    elif kind==4: # 64-bit word extreme in the tail
        pos=L-8-int(rng.integers(0,min(L-8,300))); b[pos:pos+8]=struct.pack('<q',int(rng.choice([-1,-2**63,2**63-1,0,L,L-1,2**40])))
    else:         # head region
        for _ in range(int(rng.integers(1,4))):
            b[int(rng.integers(0,min(L,400)))]=int(rng.choice(special))
    p=os.path.join(d,f'm{seed}.mcraw'); open(p,'wb').write(b)
    sys.stderr.write(f'{it} {kind}\n')
    try:
        dec=hostapi.Decoder(p)
    except hostapi.DecoderError:
        stats['open_fail']+=1; continue
    stats['open_ok']+=1
    try:
        ts=dec.get_frames(); dec.get_container_metadata(); dec.audio_sample_rate_hz(); dec.num_audio_channels()
    except hostapi.DecoderError: pass
    for mode in (False,True):
        try: dec.load_audio(mode)
        except hostapi.DecoderError: stats['audio_err']+=1
    for t in (ts[:1]+ts[-1:] if ts else []):
        try: dec.load_frame(t)
        except hostapi.DecoderError: stats['frame_err']+=1
    dec.close()
print('ok',stats)
