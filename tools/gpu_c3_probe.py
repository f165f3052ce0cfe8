import sys, os
sys.path.insert(0, os.getcwd())
import torch
from bhusie_b200 import assets, pipelines as P, uniforms as U
tex,_=assets.load_textures(); blob,_=P.load_obj_model(assets.lucy_path())
ctx=P.Context(0); ctx.set_textures(tex); ctx.upload_models(blob)
s=torch.cuda.current_stream(); rp=P.RayPipeline(ctx,3840,2160)
for cam,name in ((U.Camera(),'c3'),(U.Camera(position=(0,0,-45)),'cam45')):
  for m in (0,1):
    det=U.RayDetails(integration_method=m, model_count=1)
    for _ in range(3): rp.pass_(cam,U.BlackHole(),det,s)
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(s)
    for _ in range(8): rp.pass_(cam,U.BlackHole(),det,s)
    b.record(s); torch.cuda.synchronize()
    ms=a.elapsed_time(b)/8; st=rp.stats()
    print(name,'euler' if m==0 else 'rk', round(ms,3),'ms', round(st['ray_steps']/ms/1e6,1),'G')
