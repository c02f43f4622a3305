"""Copy the reference's scene ASSETS (JSON, .mesh, textures — data, not source) into the
git-ignored directory baseline/_ref/scenes so they travel to the GPU box with gpurun
(/root/reference does not exist there).  Run in the dev container; idempotent."""
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/scenes"
DST = os.path.join(REPO, "baseline", "_ref", "scenes")


def main():
    if not os.path.isdir(SRC):
        print(f"{SRC} not present; nothing to do (assets must already be in {DST})")
        return 0 if os.path.isfile(os.path.join(DST, "cbox.json")) else 1
    n = 0
    for root, _dirs, files in os.walk(SRC):
        rel = os.path.relpath(root, SRC)
        out = os.path.join(DST, rel) if rel != "." else DST
        os.makedirs(out, exist_ok=True)
        for f in files:
            if not f.endswith((".json", ".mesh", ".jpg")):
                continue  # OBJ/MTL importer inputs are not needed at run time
            s, d = os.path.join(root, f), os.path.join(out, f)
            if not os.path.exists(d) or os.path.getsize(d) != os.path.getsize(s):
                shutil.copyfile(s, d)
                n += 1
    print(f"assets ready in {DST} ({n} files copied)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
