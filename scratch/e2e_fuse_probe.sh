for m in 1480 256 128; do for c in 256 296 592; do echo -n "fuse_min=$m chunk=$c: "; ZIPGPU_FUSE_MIN_ROWS=$m ZIPGPU_CHUNK_ROWS=$c python scratch/e2e_zc.py | cut -c1-40; done; done
