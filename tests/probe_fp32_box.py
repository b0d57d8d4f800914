"""Why the box-filter metrics reproduce scipy's rounding instead of running in plain float32.

CPU experiment (numpy, no GPU): `_local_contrast_std` (pipeline/metrics.py:120-129) and the NIQE
var-of-var (metrics.py:195-200) with the two uniform filters evaluated (a) by scipy (float64
accumulation, float32 rounding after each axis -- the reference) and (b) in pure float32 (pairwise
tree window sums, one multiply by float32(1/size) per axis).  In smooth regions q - m^2 cancels to the
rounding noise of m and q, sqrt() amplifies it, and the metric moves by 1e-5 .. 1e-4 relative: outside
north_star's 1e-5.  Output committed as profiles/r02_fp32_box_probe.txt.

    python tests/probe_fp32_box.py
"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import numpy as np, warnings
warnings.filterwarnings('ignore')
from scipy.ndimage import uniform_filter
from mdimg_b200 import synth
from oracle import ref_metrics as omet, ref_enhancement as oenh
from test_gpu_parity import _adversarial_images

def box_f32(x, size):
    """pure float32: direct pairwise-tree window sums, multiply by float32(1/size); axis 0 then axis 1"""
    lo = size//2; hi = size - lo - 1
    def one(a, axis):
        a = np.moveaxis(a, axis, 0)
        p = np.pad(a, ((lo, hi), (0,0)), mode='symmetric')
        n = a.shape[0]
        terms = [p[k:k+n] for k in range(size)]
        while len(terms) > 1:   # pairwise tree in float32
            nxt = [terms[i] + terms[i+1] for i in range(0, len(terms)-1, 2)]
            if len(terms) % 2: nxt.append(terms[-1])
            terms = nxt
        out = (terms[0] * np.float32(1.0/size)).astype(np.float32)
        return np.moveaxis(out, 0, axis)
    return one(one(x.astype(np.float32), 0), 1)

def lcs(x, box):
    m = box(x, 7); q = box((x*x).astype(np.float32), 7)
    lv = np.maximum(q - m*m, 0)
    return float(np.std(np.sqrt(lv)))
def vov(x, box):
    m = box(x, 16); q = box((x*x).astype(np.float32), 16)
    lv = np.maximum(q - m*m, 0)
    return float(np.std(lv) / (np.mean(lv) + 1e-8))
ref_box = lambda a, s: uniform_filter(a, size=s)
ims = dict(_adversarial_images())
ct = omet.normalize_image(synth.ct_slice(1000))
ims['ct512'] = ct
enh,_ = oenh.apply_enhancements_from_params(ct, synth.plan_full())
ims['ct512_enh'] = enh.astype(np.float32)
ims['unit256'] = synth.unit_image(4000,256)
cr = omet.normalize_image(synth.radiograph(2000,600)); ims['cr600']=cr
enh2,_ = oenh.apply_enhancements_from_params(cr, synth.plan_cr()); ims['cr600_enh']=enh2.astype(np.float32)
ims['clean64']=synth.fixture_clean(); ims['noisy64']=synth.fixture_noisy(); ims['lowc64']=synth.fixture_low_contrast()
for name, im in ims.items():
    a, b = lcs(im, ref_box), lcs(im, box_f32)
    c, d = vov(im, ref_box), vov(im, box_f32)
    r1 = abs(a-b)/max(abs(a),1e-30); r2 = abs(c-d)/max(abs(c),1e-30)
    print(f"{name:12s} local_contrast ref {a:.6e} f32 {b:.6e} rel {r1:.2e} | var_of_var ref {c:.6e} f32 {d:.6e} rel {r2:.2e}")
