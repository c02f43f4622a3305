"""Command-line driver (the thin L5 layer of SURVEY §1): scene JSON + integrator JSON -> PNG / PFM.

  python tools/render.py --scene baseline/_ref/scenes/cbox.json --integrator baseline/_ref/scenes/integrator.json \
         --res 1024 1024 --spp 256 --output cbox.png
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/render.py ...   (spp split over ranks)
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aqua_engine_b200 as aq
from aqua_engine_b200 import _abi, dist as aqd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", required=True)
    ap.add_argument("--integrator", default=None, help="integrator JSON (spp, max_depth); flags below override it")
    ap.add_argument("--res", type=int, nargs=2, default=None)
    ap.add_argument("--spp", type=int, default=None)
    ap.add_argument("--max-depth", type=int, default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--type", default=None, choices=["pt", "nrc"],
                    help="override the integrator type (scenes/integrator.json says nrc)")
    ap.add_argument("--visualize-cache", action="store_true", help="nrc: show the cache at the first hit")
    ap.add_argument("--nrc-exact", action="store_true", help="nrc: bit-exact fp32 lookup instead of the tcgen05 one")
    ap.add_argument("--exposure", type=float, default=1.0)
    ap.add_argument("--output", default="out.png")
    a = ap.parse_args()
    import torch
    rank, world, local = aqd.init_from_env()
    scene = aq.Scene.load(a.scene)
    integ = aq.Integrator.load(a.integrator) if a.integrator else aq.Integrator()
    if a.spp is not None:
        integ.spp = a.spp
    if a.max_depth is not None:
        integ.max_depth = a.max_depth
    integ.seed = a.seed
    if a.type is not None:
        integ.type = a.type
    if a.visualize_cache:
        integ.visualize_cache = True
    w, h = a.res if a.res else (scene.desc.camera.res[0], scene.desc.camera.res[1])
    dr = aqd.DistRenderer(scene, local)
    film = dr.render_async(integ, w, h, nrc_exact=a.nrc_exact)
    st = dr.finish()
    torch.cuda.synchronize()
    if rank == 0:
        img = dr.r.resolve((h, w), exposure=a.exposure, d_film_ptr=film.data_ptr())
        if a.output.endswith(".pfm"):
            hf = film.cpu().numpy()
            _abi.check_host(_abi.host_lib().aq_host_write_pfm(os.fsencode(a.output), hf.ctypes.data, w, h))
        else:
            aq.write_png(a.output, img)
        s = st["ms_total"] * 1e-3
        if dr.nrc_info:
            print(f"nrc cache: {dr.nrc_info['n_records']} records in {dr.nrc_info['ms_records']:.1f} ms, "
                  f"{integ.training_iters} descent steps in {dr.nrc_info['ms_train']:.1f} ms, "
                  f"loss {dr.nrc_info['loss_first']:.3f} -> {dr.nrc_info['loss_last']:.3f}")
        print(f"{a.output}: {w}x{h}, {integ.spp} spp ({integ.type}) over {world} GPU(s), rank-0 render {st['ms_total']:.1f} ms "
              f"({(st['rays_closest'] + st['rays_shadow']) / s / 1e6:.0f} Mrays/s per GPU)")
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
