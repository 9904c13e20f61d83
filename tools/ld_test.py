import sys, torch, json
sys.path.insert(0,'/root/repo')
import rgb_experiment_b200 as P, rgb_experiment_b200.synth as S
dev=torch.device('cuda:0')
sg=S.make_named('products',device=dev,features=False); N=sg.num_nodes
g=P.Graph(sg.edge_index,N,P.LOOP_ADD_REMAINING)
d=g.dinv()
def t(fn,it=5):
    fn(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/it
for F,lds in ((24,(24,32)),(12,(12,16,32)),(47,(48,64)),(32,(32,))):
    for ld in lds:
        buf=torch.randn(N,ld,device=dev); x=buf[:,:F]
        ob=torch.empty(N,ld,device=dev); out=ob[:,:F]
        for tune in (0, 8|(1<<8)|(8<<16), 8|(1<<8)|(4<<16), 16|(1<<8)|(8<<16), 4|(2<<8)|(4<<16)):
            try:
                ms=t(lambda: P.ops.spmm_raw(g.fwd,x,None,ep=P.ops.make_epilogue(row_scale=d),out=out,tune=tune))
            except Exception as e:
                continue
            print(json.dumps({"F":F,"ld":ld,"tune":hex(tune),"ms":round(ms,4),"gteps":round(g.nnz/ms/1e6,1)}),flush=True)
