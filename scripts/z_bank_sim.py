"""Shared-memory bank-conflict simulation of the Z1 / Z2 staging stores of fb_attn_bwd.cu (build_z1 / build_z2) for candidate
(LBO, SP) layouts: average wavefronts per 32-lane 2-byte store instruction (1.0 = conflict-free).  DESIGN.md section 3."""
import itertools, collections
def conflicts(addrs):
    # addrs: list of byte addresses for 32 lanes (2-byte stores). wavefronts = max over banks of distinct 4B words
    banks=collections.defaultdict(set)
    for a in addrs: banks[(a>>2)&31].add(a>>2)
    return max(len(v) for v in banks.values())
def z1(LBO,SP,PRE,W,jfirst):
    tot=0; n=0
    for wq in range(4):
        for e in range(W):
            addrs=[]
            for lane in range(32):
                a=wq*32+lane; G=a>>3; rho=a&7
                rowbase=PRE+G*SP+rho*16+(jfirst>>3)*LBO
                r=e&7; t=7-rho+r
                off=rowbase+(t>>3)*LBO+(t&7)*2+(e>>3)*LBO
                addrs.append(off)
            tot+=conflicts(addrs); n+=1
    return tot/n
def z2(LBO,SP,PRE,W,jfirst):
    tot=0;n=0
    for wq in range(4):
        for e in range(W):
            addrs=[]
            for lane in range(32):
                a=wq*32+lane
                base=PRE+(jfirst>>3)*SP
                r=e&7; y=a+7-r
                off=base+r*16+(y>>3)*LBO+(y&7)*2+(e>>3)*SP
                addrs.append(off)
            tot+=conflicts(addrs); n+=1
    return tot/n
for LBO in (128,144,160,272):
    for SP in (4000,4016,4032,4048,4064,4080):
        PRE=15*LBO
        print(LBO,SP,"z1 %.2f z2 %.2f"%(z1(LBO,SP,PRE,64,0),z2(LBO,SP,PRE,64,64)), "bytes",PRE+15*SP+32*LBO)
print("---- search")
best=[]
for LBO in range(128,321,16):
    for SP in range(((17*LBO+15)//16)*16, 4400, 16):
        PRE=15*LBO
        a=z1(LBO,SP,PRE,64,0); b=z2(LBO,SP,PRE,64,64)
        best.append((a+b,a,b,LBO,SP,PRE+15*SP+32*LBO))
best.sort()
for x in best[:15]: print(x)
print("---- constrained search")
best=[]
for LBO in (128,144,160):
    lo=int(31.25*LBO); lo=(lo+15)//16*16
    for SP in range(lo, lo+400, 16):
        PRE=15*LBO
        a=z1(LBO,SP,PRE,64,0); b=z2(LBO,SP,PRE,64,64)
        best.append((a+b,a,b,LBO,SP,PRE+15*SP+32*LBO))
best.sort()
for x in best[:12]: print(x)
