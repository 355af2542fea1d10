import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import amp_sparc_spatialmodulation_b200 as pkg
from conftest import load_golden, config_from_meta
from test_gpu_parity import run_bamp_golden
from parity_utils import rel_err
for name in ['bamp_c1','bamp_c2']:
  for mode in [dict(kernel='generic',exp='f32'), dict(kernel='fast',exp='f32')]:
    try:
        g,out=run_bamp_golden(name,**mode)
    except Exception as e:
        print(name,mode,'ERR',e); continue
    for j,key in enumerate(['tau','varm','mse']):
        r=rel_err(out['traj'][:,:,j],g[key])
        i=np.unravel_index(np.argmax(r),r.shape)
        print(name,mode['kernel'],key,'worst rel',r.max(),'at frame,it',i,'got',out['traj'][i[0],i[1],j],'want',g[key][i],'snr',g['snr_db'][i[0]],'median',np.median(r), 'early max', r[:,:2].max())
    print('   iters equal frac',(out['iters']==g['iters']).mean(),'max|dx|',np.abs(out['xmmse']-g['xmmse']).max())
    for snr,have,want,_ in out['counters']:
        bad={k:(have[k],want[k]) for k in want if k in have and isinstance(want[k],int) and have[k]!=want[k] and k not in('iters','nan_frames')}
        print('   snr',snr,'count mismatches',bad)
