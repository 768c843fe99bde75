"""Model of the barrier protocol of the round-2 candidate kernel (conv_tma_fast_kernel in
confignet_b200/csrc/experiments/round2_probes.cu): producer (TMA + bulk copy into a ring of S stages), a_small builders
(tensor-memory stage per ring slot), MMA issue (ping-pong accumulators, chunks of CH k-blocks), promotion / epilogue.
Each agent is a generator that mirrors the kernel's loop, stage / phase arithmetic and mbarrier parity waits; a random
scheduler interleaves them and the model asserts the hazards the barriers exist to prevent (a stage overwritten before its
MMAs retired, an a_small stage rebuilt while still being read, an accumulator restarted before it was promoted, a chunk
promoted twice or not at all) and that no interleaving deadlocks.  Pure Python; run by
tests/test_host_cpu.py::test_candidate_kernel_protocol_model.

    python scripts/sim_candidate_protocol.py
"""
import itertools, random
class Bar:
    def __init__(s, count): s.count=count; s.pending=count; s.done=0
    def arrive(s):
        s.pending-=1
        if s.pending==0: s.done+=1; s.pending=s.count
    def passed(s, parity): return (s.done & 1) != parity
def run(S, num_kb, tiles, CH=8, seed=0, producer_waits_empty=True):
    rnd=random.Random(seed)
    kbs = list(num_kb) if isinstance(num_kb, (list, tuple)) else [num_kb]      # per item, cycled: phases differ in tap count
    full=[Bar(1) for _ in range(S)]; small=[Bar(1) for _ in range(S)]; empty=[Bar(1) for _ in range(S)]
    accfull=[Bar(1),Bar(1)]; accempty=[Bar(1),Bar(1)]
    stage_owner=[None]*S      # (tile,kb) currently loaded in smem stage
    tmem_small=[None]*S
    acc=[None,None]           # chunk id accumulated, state
    log=[]
    def producer():
        s=0;ph=0;git=0
        for t in range(tiles):
            for kb in range(kbs[t % len(kbs)]):
                if git>=S and producer_waits_empty:
                    while not empty[s].passed(ph^1): yield
                assert stage_owner[s] is None or stage_owner[s][2]=='consumed', ('overwrite smem', t,kb,s,stage_owner[s])
                stage_owner[s]=[t,kb,'loaded']; full[s].arrive()
                s+=1; git+=1
                if s==S: s=0; ph^=1
                yield
    def builder():
        s=0;ph=0;git=0
        for t in range(tiles):
            for kb in range(kbs[t % len(kbs)]):
                while not full[s].passed(ph): yield
                if git>=S:
                    while not empty[s].passed(ph^1): yield
                assert stage_owner[s][:2]==[t,kb], ('builder reads wrong stage', t,kb,stage_owner[s])
                assert tmem_small[s] is None or tmem_small[s][2]=='consumed', ('overwrite tmem small', t,kb)
                tmem_small[s]=[t,kb,'built']; small[s].arrive()
                s+=1; git+=1
                if s==S: s=0; ph^=1
                yield
    def mma():
        s=0;ph=0;b=0;inchunk=0;c=0
        for t in range(tiles):
            num_kb = kbs[t % len(kbs)]
            for kb in range(num_kb):
                chunk_first = inchunk==0; chunk_last = inchunk==CH-1 or kb==num_kb-1
                if chunk_first and c>=2:
                    while not accempty[b].passed(((c>>1)-1)&1): yield
                while not full[s].passed(ph): yield
                while not small[s].passed(ph): yield
                assert stage_owner[s][:2]==[t,kb] and stage_owner[s][2]=='loaded', ('mma wrong smem', t,kb,stage_owner[s])
                assert tmem_small[s][:2]==[t,kb] and tmem_small[s][2]=='built', ('mma wrong tmem', t,kb,tmem_small[s])
                if chunk_first:
                    assert acc[b] is None or acc[b][1]=='read', ('acc overwritten before read', t,kb,b,acc[b])
                    acc[b]=[c,'accumulating',0]
                assert acc[b][0]==c
                acc[b][2]+=1
                stage_owner[s][2]='consumed'; tmem_small[s][2]='consumed'; empty[s].arrive()
                if chunk_last: acc[b][1]='full'; accfull[b].arrive()
                s+=1
                if s==S: s=0; ph^=1
                inchunk+=1
                if inchunk==CH or kb==num_kb-1: inchunk=0; c+=1; b^=1
                yield
    def epilogue():
        b=0;c=0
        for t in range(tiles):
            num_kb = kbs[t % len(kbs)]
            nchunks=(num_kb+CH-1)//CH; total=0
            for ch in range(nchunks):
                while not accfull[b].passed((c>>1)&1): yield
                assert acc[b][0]==c and acc[b][1]=='full', ('epilogue reads wrong acc', t,ch,c,acc[b])
                total+=acc[b][2]; acc[b][1]='read'; accempty[b].arrive()
                b^=1; c+=1
                yield
            assert total==num_kb, (t,total)
            log.append(t)
    agents=[producer(),builder(),mma(),epilogue()]
    alive=[True]*4; idle=0
    while any(alive):
        i=rnd.randrange(4)
        if not alive[i]: continue
        before=(tuple(b.done for b in full+small+empty+accfull+accempty), len(log))
        try: next(agents[i])
        except StopIteration: alive[i]=False
        after=(tuple(b.done for b in full+small+empty+accfull+accempty), len(log))
        idle = idle+1 if before==after else 0
        assert idle<20000, 'deadlock'
    assert log==list(range(tiles))
def main():
    for S, num_kb, tiles, seed in itertools.product((2, 3, 4, 6), (1, 7, 8, 9, 17, 18, 72), (1, 2, 5), (0, 1, 2)):
        run(S, num_kb, tiles, seed=seed)
    for S, seed in itertools.product((2, 3, 6), (0, 1)):
        run(S, (12, 6, 6, 3), 9, seed=seed)                     # the four parity phases of a stride-2 input gradient, 96 channels
        run(S, (64,), 4, seed=seed)
    print("protocol ok")


if __name__ == "__main__":
    main()
