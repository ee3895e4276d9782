import numpy as np, sys, importlib
sys.path.insert(0,'/root/repo')
from tests import helpers as H
from oracle import orc
cozk=importlib.import_module("co-zkvms_b200")
ctx=cozk.Context()
rng=np.random.default_rng(5)
P=H.P
vals=[0,1,2,P-1,P-2,(1<<253),P>>1,(1<<29)-1,(1<<232)]+[int.from_bytes(rng.bytes(32),'little')%P for _ in range(20000)]
a=np.stack([H.le32(v) for v in vals]); b=np.roll(a,5,axis=0)
for op,o in (("f29_mul","mul"),("f29_add","add"),("f29_sub","sub")):
    want=orc.field_op("fq",o,a.view(np.uint64),b.view(np.uint64)).view(np.uint8)
    print(op,(ctx.field_op(op,a,b)==want).all())
print("sqr",(ctx.field_op("f29_sqr",a)==orc.field_op("fq","sqr",a.view(np.uint64)).view(np.uint8)).all())
print("inv",(ctx.field_op("f29_inv",a[:600])==orc.field_op("fq","inv",a[:600].view(np.uint64)).view(np.uint8)).all())
n=4000
pts=np.zeros((n,72),np.uint8); pts[:,:64]=orc.gen_bases(2,n)
other=np.roll(pts,1,axis=0)
other[::7]=pts[::7]; other[3::11]=orc.g1_op("neg",pts[3::11]); pts[5::13]=H.point_wire(None)
print("madd3",(ctx.g1_op("f29_madd3",pts,other)==orc.g1_op("add",pts,other)).all())
