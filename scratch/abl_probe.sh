for a in 17 49 81 0 32 64; do echo -n "abl=$a: "; ZIPGPU_ABL=$a python scratch/enc_only.py; done
