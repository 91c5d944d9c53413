"""Long random walks of the three CPU-side parity checks, beyond the seeds the test-suite runs every time (needs
/root/reference or a prebuilt oracle/_ref; no GPU):
  abi      product vs reference: random encoder call sequences, decoder observations on cut / damaged streams
           (tests/test_abi_differential.py with other seeds)
  kernels  the kernels' codec code compiled for the host vs the oracle, random parameter sets (tests/test_hostemu.py)
  oracle   the oracle vs the unmodified reference, random parameter sets (tests/test_oracle.py)
  damaged  damaged scans (bit flips, truncation, spliced bytes): the kernels' codec code on the host and the oracle, each
           against the unmodified reference as the arbiter; prints how often each of them accepts / rejects / agrees
usage: python tools/parity_hunt.py abi|kernels|oracle|damaged
End of round 1: abi 388 + 144 seeds, kernels 6080 cases, oracle 1500 cases -- no discrepancy; damaged: DESIGN.md section 8."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def hunt_abi():
    import random, sys, traceback
    from charls_b200 import capi
    from tests.support import reference_library
    import tests.test_abi_differential as T
    product = capi.default_library(); reference = reference_library()
    bad = 0
    for seed in range(12, 400):
        rng = random.Random(1000 + seed)
        try:
            for _ in range(25):
                pair = T.EncoderPair(product, reference)
                try:
                    T.random_encoder_sequence(pair, rng, rng.randrange(4, 40))
                finally:
                    pair.close()
        except AssertionError as e:
            bad += 1
            print("ENC seed", seed, str(e)[:300]); 
            if bad > 5: break
    print("encoder hunt done", bad)
    bad = 0
    for seed in range(6, 150):
        rng = random.Random(77 + seed)
        try:
            for stream in T.header_streams(reference, rng, 20):
                assert T.decoder_observations(product, stream) == T.decoder_observations(reference, stream)
                for _ in range(6):
                    cut = rng.randrange(0, len(stream))
                    assert T.decoder_observations(product, stream, cut) == T.decoder_observations(reference, stream, cut), cut
                first_scan = stream.find(b"\xff\xda")
                limit = first_scan if first_scan > 0 else len(stream)
                for _ in range(6):
                    damaged = bytearray(stream)
                    damaged[rng.randrange(2, limit)] = rng.randrange(256)
                    a = T.decoder_observations(product, bytes(damaged)); b = T.decoder_observations(reference, bytes(damaged))
                    if a != b:
                        bad += 1
                        print("DEC seed", seed, [ (x,y) for x,y in zip(a,b) if x!=y][:3], bytes(damaged)[:80].hex())
                        if bad > 8: raise SystemExit
        except AssertionError as e:
            bad += 1; print("DEC seed", seed, "assert", str(e)[:200])
            if bad > 8: break
    print("decoder hunt done", bad)


def hunt_kernels():
    import random, sys
    import numpy as np
    from tests.support import oracle, s_smooth, s_noise, s_mixed
    from tests.hostemu_lib import HostEmu
    import tests.test_hostemu as T
    o=oracle(); he=HostEmu()
    bad=0; n=0
    for seed in range(8, 160):
        rng = random.Random(4000 + seed)
        for _ in range(40):
            bits = rng.choice([2, 3, 5, 7, 8, 8, 9, 10, 12, 12, 15, 16, 16])
            maxval = (1 << bits) - 1
            cc = rng.choice([1, 1, 1, 2, 3, 3, 4])
            ilv = 0 if cc == 1 else rng.choice([1, 2])
            near = rng.choice([0, 0, 0, 1, 2, 3, min(255, maxval // 2)])
            near = min(near, maxval // 2, 255)
            xf = rng.choice([0, 1, 2, 3]) if (cc == 3 and near == 0 and bits in (8, 16)) else 0
            ri = rng.choice([1, 1, 1, 0, 2, 5])
            w, h = rng.choice([1, 2, 3, 17, 64, 65, 127, 200, 301, 1025]), rng.choice([1, 2, 5, 9])
            pc = None
            if rng.random() < 0.35:
                t1 = rng.randint(near + 1, maxval); t2 = rng.randint(t1, maxval); t3 = rng.randint(t2, maxval)
                pc = (0, t1, t2, t3, rng.choice([3, 4, 31, 64, 255]))
            gen = rng.choice([s_smooth, s_noise, s_mixed])
            img = gen(h, w, bits, cc, seed=rng.randrange(1 << 30), layout="interleaved") if cc > 1 else gen(h, w, bits, seed=rng.randrange(1 << 30))
            n+=1
            try:
                T.check_scan(o, he, img, bits, cc, near, ilv, xf, ri, pc)
            except AssertionError as e:
                bad+=1; print("seed",seed,str(e)[:200])
                if bad>5: raise SystemExit
    print("done", n, "cases", bad, "bad")


def hunt_oracle():
    import random, sys
    import numpy as np
    from charls_b200 import codec
    from tests.support import oracle, reference_library, s_smooth, s_noise, s_mixed
    from tests import jlsio
    o=oracle(); ref=reference_library()
    def payloads(stream):
        s=jlsio.parse(stream); return [stream[sc.data_offset:sc.data_end] for sc in s.scans]
    bad=0;n=0
    rng=random.Random(99)
    for _ in range(1500):
        bits=rng.choice([2,3,4,5,6,7,8,8,9,10,11,12,13,14,15,16])
        maxval=(1<<bits)-1
        cc=rng.choice([1,1,2,3,3,4])
        ilv=rng.choice([0,1,2]) if cc>1 else 0
        near=min(rng.choice([0,0,1,2,3,7,maxval//2]),maxval//2,255)
        xf=rng.choice([0,1,2,3]) if (cc==3 and near==0 and bits in (8,16) and ilv!=0) else 0
        w,h=rng.choice([1,2,3,9,33,64,130]),rng.choice([1,2,3,7,12])
        pc=None
        if rng.random()<0.3:
            t1=rng.randint(near+1,maxval); t2=rng.randint(t1,maxval); t3=rng.randint(t2,maxval)
            pc=(maxval,t1,t2,t3,rng.choice([3,4,31,64,255]))
        gen=rng.choice([s_smooth,s_noise,s_mixed])
        img=gen(h,w,bits,cc,seed=rng.randrange(1<<30),layout="planar" if ilv==0 else "interleaved") if cc>1 else gen(h,w,bits,seed=rng.randrange(1<<30))
        n+=1
        try:
            a=codec.encode(img,bits,near_lossless=near,interleave_mode=ilv,color_transformation=xf,preset=pc,lib=ref)
            b=o.encode_image(img,bits,near=near,ilv=ilv,xform=xf,pc=pc)
            assert payloads(a)==payloads(b),"encode"
            for ri in (1,rng.choice([2,3,5])):
                c=o.encode_image(img,bits,near=near,ilv=ilv,xform=xf,pc=pc,ri=ri)
                want,_,_=codec.decode(c,lib=ref); got,_=o.decode_image(c)
                assert np.array_equal(got,want),("decode",ri)
        except Exception as e:
            bad+=1; print((img.shape,bits,near,ilv,xf,pc),repr(e)[:200])
            if bad>6: break
    print("done",n,bad)


def hunt_damaged():
    import numpy as np
    from charls_b200 import codec
    from charls_b200.capi import CharlsError
    from tests.support import oracle, reference_library, s_mixed, s_noise, s_smooth
    from tests.hostemu_lib import HostEmu
    from tests import jlsio
    o=oracle(); he=HostEmu(); ref=reference_library()
    rng=np.random.default_rng(3)
    stats={}
    for bits,cc,ilv,near,ri in ((8,1,0,0,0),(8,1,0,0,1),(12,1,0,2,3),(16,3,2,0,0),(8,3,1,0,2),(16,1,0,0,1),(5,4,2,1,0),(8,3,2,0,1),(2,1,0,0,0)):
        for gen in (s_mixed,s_smooth,s_noise):
            img=gen(7,90,bits,cc,seed=bits+cc,layout="interleaved") if cc>1 else gen(7,90,bits,seed=bits)
            sp=o.params(90,7,bits,cc,near,ilv,0,None,ri); good=o.encode_scan(sp,img); hp=he.params(sp)
            whole=o.encode_image(img,bits,near=near,ilv=ilv,ri=ri)
            parsed=jlsio.parse(whole); sc=parsed.scans[0]
            assert whole[sc.data_offset:sc.data_end]==good
            for trial in range(60):
                data=bytearray(good)
                kind=trial%3
                if kind==0:
                    for _ in range(1+trial%4):
                        i=int(rng.integers(0,len(data))); data[i]^=1<<int(rng.integers(0,8))
                elif kind==1:
                    cut=int(rng.integers(1,len(data))); data=data[:cut]
                else:
                    i=int(rng.integers(0,len(data))); data[i:i+int(rng.integers(1,6))]=bytes(rng.integers(0,256,size=int(rng.integers(0,5)),dtype=np.uint8))
                data=bytes(data)
                stream=whole[:sc.data_offset]+data+b"\xff\xd9"
                try:
                    want,_,_=codec.decode(stream,lib=ref); r=0
                except CharlsError as e:
                    r=e.errc
                want_o=np.zeros_like(img); n1=o.decode_scan(sp,data+b"\xff\xd9",want_o)
                got=np.zeros_like(img); n2=he.decode(hp,data+b"\xff\xd9",got,False)
                key=("ref ok" if r==0 else "ref err", "oracle ok" if n1>=0 else "oracle err", "ours ok" if n2>=0 else "ours err")
                same = (r==0 and n2>=0 and np.array_equal(got.reshape(want.shape) if got.size==want.size else got, want))
                same_o = (r==0 and n1>=0 and np.array_equal(want_o.reshape(want.shape), want))
                stats[key+(("ours==ref" if same else "ours!=ref") if r==0 and n2>=0 else "", ("oracle==ref" if same_o else "oracle!=ref") if r==0 and n1>=0 else "")]=stats.get(key+(("ours==ref" if same else "ours!=ref") if r==0 and n2>=0 else "", ("oracle==ref" if same_o else "oracle!=ref") if r==0 and n1>=0 else ""),0)+1
    for k,v in sorted(stats.items(), key=lambda x:-x[1]): print(v,k)


if __name__ == "__main__":
    {"abi": hunt_abi, "kernels": hunt_kernels, "oracle": hunt_oracle, "damaged": hunt_damaged}[sys.argv[1]]()
