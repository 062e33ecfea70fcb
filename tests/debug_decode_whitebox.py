# TEST INFRASTRUCTURE (debugging aid, uses the oracle helpers): white-box comparison of the decoder's internal buffers with a Python restatement.
import os, sys, json
import numpy as np
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
from agmv_testlib import *
from test_gpu_parity import _chunk_ranges
golden = json.load(open(os.path.join(GOLDEN_DIR, "golden.json")))
FILL,NORM,COPY=0x4E,0x2F,0x5E
def expand(data, start, usize, csize):
    bitp=start*8; nb=csize*8; bits=0; out=bytearray()
    def rd(n):
        nonlocal bitp
        v=0
        for k in range(n):
            byte=data[bitp>>3] if (bitp>>3)<len(data) else 0
            v|=((byte>>(bitp&7))&1)<<k; bitp+=1
        return v
    while bits<nb and len(out)<usize:
        f=rd(1); bits+=1
        if f: out.append(rd(8)); bits+=8
        else:
            off=rd(16); l=rd(4); bits+=20; p=len(out)
            for i in range(l):
                s=p-off+i
                if 0<=s<len(out) and off>=1: out.append(out[s])
    return out
def emul(data, trace=None):
    W=int.from_bytes(data[8:12],'little'); H=int.from_bytes(data[12:16],'little'); nfr=int.from_bytes(data[4:8],'little')
    dual = data[17] in (1,3)
    pal=[(data[38+3*i]<<16)|(data[39+3*i]<<8)|data[40+3*i] for i in range(512 if dual else 256)]+[0]*256
    B=(W//4)*(H//4); bw=W//4
    img=np.zeros((H,W),np.uint32); ifr=img.copy(); persist=bytearray(2*W*H+64)
    frames=[]
    chunks=_chunk_ranges(data)
    for k,(start,cs) in enumerate(chunks[:nfr]):
        us=int.from_bytes(data[start-8:start-4],'little')
        e=expand(data,start,us,cs); bpos=len(e)
        buf=bytearray(persist); buf[:bpos]=e   # virtual buffer
        persist[:bpos]=e
        limit=max(bpos-72,0)
        # step codes
        def code(p):
            b=buf[p]
            if b==COPY: return 1,1
            if b==FILL: return 2+(1 if dual and (buf[p+1]&0x7f)==127 else 0),1
            if b==NORM:
                q=p+1
                for _ in range(16): q+= 2 if dual and (buf[q]&0x7f)==127 else 1
                return q-p,1
            return 1,0
        rec=[None]*B
        p=0;cum=0
        while p<limit:
            st,w=code(p)
            if w and cum<B:
                t = 2 if st==1 else (1 if st<=3 else 3)
                rec[cum]=(p+1,t)
            cum+=w; p+=st
        # tail walk
        bp=p; b=min(cum,B); invalid=False
        def isflag(x): return x in (FILL,NORM,COPY)
        while b<B:
            if bp>bpos: break
            fl=buf[bp]; bp+=1; esc=False
            while not isflag(fl):
                fl=buf[bp]; bp+=1
                if bp>bpos: esc=True; break
            if not isflag(fl): invalid=True
            if fl==FILL:
                at=bp; c=buf[bp]; bp+=1
                if dual and (c&0x7f)==127: bp+=1
                if bp>bpos: rec[b]=None; b+=1; break
                rec[b]=(at,1)
            elif fl==COPY: rec[b]=(bp,2)
            else:
                at=bp; wrote=False; kk=0
                while kk<16:
                    c=buf[bp]; bp+=1
                    if dual and (c&0x7f)==127: bp+=1
                    if bp>bpos or invalid:
                        esc=True; invalid=False; kk=(kk|3)+1; continue
                    wrote=True; kk+=1
                rec[b]=(at,3) if wrote else None
            if esc: b+=1; break
            b+=1
        # reconstruct
        def color(bp):
            c=buf[bp]; bp+=1
            if not dual: return pal[c],bp
            base=256 if c&0x80 else 0
            if (c&0x7f)<127: return pal[base+(c&0x7f)],bp
            v=pal[base+buf[bp]]; return v,bp+1
        new=img.copy()
        for b in range(B):
            x=(b%bw)*4; y=(b//bw)*4
            r=rec[b]
            if r is None: continue
            at,t=r
            if t==2: new[y:y+4,x:x+4]=ifr[y:y+4,x:x+4]
            elif t==1:
                col,_=color(at)
                if b==B-1: col=new[y+1,x-1]
                new[y:y+4,x:x+4]=col
            else:
                bp2=at
                for i in range(16):
                    col,bp2=color(bp2)
                    if bp2<=bpos: new[y+i//4,x+i%4]=col
        if trace is not None: trace.append((bpos, [ (r[0]<<2|r[1]) if r else 0xFFFFFFFF for r in rec], bytes(e)))
        img=new
        if k%4==0: ifr=img.copy()
        frames.append(img.copy())
    return np.stack(frames)

import libagmv_b200
ctx = libagmv_b200.Context(0)
for name in ["syn96x80_III_LOW", "gba240_GBA_I_LOW"]:
    g = golden["encode"][name]
    clean = open(os.path.join(GOLDEN_DIR, g["file"]), "rb").read()
    rng = np.random.default_rng(4242)
    data = bytearray(clean)
    for ci, (start, cs) in enumerate(_chunk_ranges(clean)):
        if cs < 8 or rng.random() < 0.3: continue
        for _ in range(int(rng.integers(1, 4))):
            at = start + int(rng.integers(0, cs)); data[at] = int(rng.integers(0, 256))
        if rng.random() < 0.3:
            newcs = int(rng.integers(cs // 2, cs)); data[start - 4:start] = newcs.to_bytes(4, "little")
    data=bytes(data)
    tr=[]
    exp=emul(data,tr)
    got=ctx.decode_all(data)
    n,H,W=exp.shape; B=(W//4)*(H//4)
    bpos=ctx.test_peek(0,np.uint32,n); recs=ctx.test_peek(1,np.uint32,n*B).reshape(n,B); stale=ctx.test_peek(3,np.uint8,n*4).reshape(n,4)
    print('gpu bpos ',list(bpos)); print('emul bpos',[t[0] for t in tr]); print('usize', [int.from_bytes(data[st-8:st-4],'little') for st,cs in _chunk_ranges(data)][:n]); print('csize', [cs for st,cs in _chunk_ranges(data)][:n]); print('clean cs', [cs for st,cs in _chunk_ranges(clean)][:n])
    for k in range(n):
        eb,er,ee=tr[k]
        if bpos[k]!=eb: print(name,'frame',k,'bpos gpu',bpos[k],'emul',eb); break
        er=np.array(er,dtype=np.uint32)
        if not np.array_equal(recs[k],er):
            d=np.argwhere(recs[k]!=er).ravel(); print(name,'frame',k,'recs differ at blocks',d[:8],'gpu',[hex(x) for x in recs[k][d[:4]]],'emul',[hex(x) for x in er[d[:4]]],'bpos',eb,'limit',max(eb-72,0)); break
        if not np.array_equal(got[k],exp[k]): print(name,'frame',k,'pixels differ but recs/bpos equal; stale',stale[k]); break
    else: print(name,'all equal')
