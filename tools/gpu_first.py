import sys, time
sys.path.insert(0,'.')
import numpy as np
from tilecoderaytracer_b200 import api
from oracle import oracle_py as O
import __graft_entry__ as G
G.smoke()
ctx=api.Context([0])
for name,W,H,D in [("default",500,504,50),("synth1024",200,160,50),("synth256",200,160,10),("random:1:40",200,160,8),("random:7:120",160,120,12),("random:3:300",160,120,20),("two_mirrors",100,80,50)]:
    cam=api.Camera(); sc=api.Scene().build(name,cam)
    ctx.upload(sc,cam); p=api.default_params(W,H,D)
    img,st=ctx.render(p)
    ref,cnt=O.render(sc.flatten(),cam.export(),p)
    bad=int((img.view(np.uint32)!=ref.view(np.uint32)).any(-1).sum())
    txt_ok = ctx.format_txt()==O.format_txt(ref)
    print(name,"mismatch px",bad,"of",W*H,"txt_ok",txt_ok,"ms",st.render_ms,"rays",st.rays,cnt["rays_primary"]+cnt["rays_shadow"]+cnt["rays_reflect"], flush=True)
# perf: big configs
for name,W,H,D in [("default",1920,1080,5),("default",3840,2160,50),("synth1024",3840,2160,50),("synth256",7680,4320,10),("two_mirrors",1920,1080,50)]:
    cam=api.Camera(); sc=api.Scene().build(name,cam)
    ctx.upload(sc,cam); p=api.default_params(W,H,D)
    for i in range(3):
        st=ctx.render_device(p)
    print(name,W,H,D,"ms",st.render_ms[0],"Mrays/s",st.rays/st.render_ms[0]/1e3,"rays",st.rays,flush=True)
