cd $GRAFT_REPO_ROOT
cat > /tmp/t.py <<'P'
import sys; sys.path.insert(0,'.')
import coolchic_b200, numpy as np
from coolchic_b200.bitstream.decode import decode_video
fr = decode_video('tests/golden/kodim14.cool')['0']
ref = np.load('tests/golden/kodim14_image_u8.npz')['image']
got = np.round(fr.data[0].numpy()*255).astype(np.uint8).transpose(1,2,0)
print("diff", int((got!=ref).sum()))
P
for m in 0 1 3; do echo "== CCD_TMA_MODE=$m"; CCD_TMA_MODE=$m timeout 60 python /tmp/t.py 2>&1 | tail -2; done
echo "== sanitizer (mode 1)"; CCD_TMA_MODE=1 timeout 200 compute-sanitizer --tool memcheck python /tmp/t.py 2>&1 | grep -v "^$" | head -30
