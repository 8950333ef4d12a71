"""Discrete-event model of the barrier protocol of v2h_train_tc_kernel (debugging aid): producer, polling issuer, in-order tensor
pipe with commit semantics, two epilogue groups; mbarriers with (count, phase) and parity waits."""
import sys, random

class Bar:
    def __init__(s, count): s.count=count; s.pending=count; s.phase=0
    def arrive(s):
        s.pending-=1
        assert s.pending>=0
        if s.pending==0: s.phase+=1; s.pending=s.count
    def test(s, parity):  # True iff the phase with this parity has completed: current phase parity != parity
        return (s.phase & 1) != parity

def run(my_units, XS, npsets, STAT=True, seed=0, steps=200000):
    rnd=random.Random(seed)
    x_full=[Bar(1) for _ in range(8)]; x_free=[Bar(1) for _ in range(8)]
    d1_full=[Bar(1) for _ in range(4)]; d1_free=[Bar(1) for _ in range(4)]
    p_full=[Bar(1) for _ in range(2)]; p_free=[Bar(1) for _ in range(2)]
    pipe=[]   # tensor pipe queue: ('mma', dur) or ('commit', bar)
    st=dict(prod_t=0, g1=0, g2=0)
    # epilogue groups: state machines
    ep=[dict(t=pr, stage='wait_d1', busy=0) for pr in range(2)]
    npm=npsets-1
    done_units=[False]*my_units
    time=0
    pipe_busy=0
    for it in range(steps):
        time+=1
        # tensor pipe
        if pipe_busy>0: pipe_busy-=1
        while pipe_busy==0 and pipe:
            kind,arg=pipe[0]
            if kind=='mma': pipe_busy=arg; pipe.pop(0); break
            else: arg.arrive(); pipe.pop(0)
        # producer
        if st['prod_t']<my_units and rnd.random()<0.3:
            t=st['prod_t']; stage=t%XS; use=t//XS
            if use==0 or x_free[stage].test((use-1)&1):
                x_full[stage].arrive(); st['prod_t']+=1
        # issuer (polling)
        if rnd.random()<0.5:
            g2_total=my_units if STAT else 0
            if STAT and st['g2']<st['g1']:
                g2=st['g2']; set_i=g2&npm; use=(g2>>1) if npsets==2 else g2
                if p_full[set_i].test(use&1):
                    pipe.append(('mma', 8)); pipe.append(('commit', x_free[g2%XS])); pipe.append(('commit', p_free[set_i if npsets==2 else (g2+1)&1])); st['g2']+=1
            if st['g1']<my_units:
                t=st['g1']; stage=t%XS; buf=t&3
                if x_full[stage].test((t//XS)&1) and (t<4 or d1_free[buf].test(((t>>2)-1)&1)):
                    pipe.append(('mma', 4)); pipe.append(('commit', d1_full[buf]))
                    if not STAT: pipe.append(('commit', x_free[stage]))
                    st['g1']+=1
        # epilogue groups
        for pr in range(2):
            e=ep[pr]
            if e['t']>=my_units: continue
            if e['busy']>0: e['busy']-=1; continue
            t=e['t']; buf=t&3
            if e['stage']=='wait_d1':
                if d1_full[buf].test((t>>2)&1):
                    d1_free[buf].arrive(); e['stage']='compute'; e['busy']=rnd.randint(5,40)
            elif e['stage']=='compute':
                if not STAT:
                    done_units[t]=True; e['t']+=2; e['stage']='wait_d1'; continue
                set_i=t&npm; use=(t>>1) if npsets==2 else t
                if npsets==2:
                    if use>=1 and not p_free[set_i].test((use-1)&1): continue
                else:
                    if t>=1 and not p_free[t&1].test(((t-1)>>1)&1): continue
                p_full[set_i].arrive(); done_units[t]=True; e['t']+=2; e['stage']='wait_d1'
        if all(done_units) and st['g2']>=(my_units if STAT else 0) and not pipe and pipe_busy==0:
            return time
    return None

if __name__=='__main__':
    bad=0
    for XS in (3,4,6,8):
        for npsets in (1,2):
            for units in (1,2,3,5,8,13,29):
                for seed in range(20):
                    r=run(units,XS,npsets,True,seed)
                    if r is None:
                        bad+=1; print('DEADLOCK XS',XS,'npsets',npsets,'units',units,'seed',seed); break
    print('bad',bad)
