import sys; sys.path.insert(0,'/root/repo')
import numpy as np
import dataset_pipeline_b200 as b2
from dataset_pipeline_b200 import registration as R
from dataset_pipeline_b200.synth import reg_scene
from oracle import oracle
sc = reg_scene.make_scene(num_images=2, width=320, height=240, fx=260.0, camera_model=5, num_scales=3, base_radius=0.004)
area = 320*240//4
g = b2.Registration(R.default_params(max_initial_image_area_in_pixels=area)); o = oracle.Registration(oracle.reg_default_params(max_initial_image_area_in_pixels=area))
reg_scene.load_into(g, sc); reg_scene.load_into(o, sc)
g.set_image_scale(0); o.set_image_scale(0)
g.CreateObservationsForAllImages(1); o.create_observations(1)
for im in range(2):
    for ps in range(3):
        gi,gx,gy,gs,gn = g.observations(im,ps); oi,ox,oy,os_,on = o.observations(im,ps)
        common, ga, oa = np.intersect1d(gi, oi, return_indices=True)
        print(im, ps, len(gi), len(oi), len(common), "dx", np.abs(gx[ga]-ox[oa]).max(initial=0), "ds", np.abs(gs[ga]-os_[oa]).max(initial=0), "nb diff", int((gn[ga]!=on[oa]).sum()))
g.ColorOptimizerApply(); o.color_update()
Hg,bg,sg,cg = g.accumulate(); Ho,bo,so,co = o.accumulate()
print("H rel", np.abs(Hg-Ho).max()/np.abs(Ho).max(), "b rel", np.abs(bg-bo).max()/np.abs(bo).max(), "cost", cg, co, abs(cg-co)/co, sg, so)
D=np.abs(Hg-Ho)/ (np.sqrt(np.outer(np.diag(Ho),np.diag(Ho)))+1e-30)
print("scaled H err max", D.max())
