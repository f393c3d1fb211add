"""Run under torchrun with 2+ GPUs (scripts/gpu_multi.sh): vocab-parallel step == replicated data-parallel step,
sharded top-k == local top-k.  Prints 'multi-gpu check ok' on rank 0."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    import ac_tsr_b200 as A
    from oracle import acsr_oracle as O
    from test_gpu_model import make_config, DS
    cfg = O.default_cfg(hidden_dropout_prob=0.0, attn_dropout_prob=0.0)
    V, B, L, k = 5003, 64, 50, 50
    params = O.init_params(cfg, V, seed=11)
    seq, ln, pos = O.synth_batch(B * world, L, V, seed=12)
    sl = slice(rank * B, (rank + 1) * B)
    g = torch.Generator().manual_seed(100 + rank)
    noise = {(l, 'noise'): torch.randn(B, cfg['n_heads'], L, L, generator=g).to(dev) for l in range(cfg['n_layers'])}
    res = []
    for vocab_parallel in (False, True):
        config = make_config(A, cfg, cuda_graph=False, device=dev)
        model = A.ACSASRec(config, DS(V)).to(dev)
        model.load_state_dict({kk: v.to(dev) for kk, v in params.items()})
        model._debug_rand = noise
        trainer = A.ACSASRecTrainer(config, model)
        trainer.enable_data_parallel(vocab_parallel=vocab_parallel)
        model.train()
        inter = A.Interaction({'item_id_list': seq[sl].to(dev), 'item_length': ln[sl].to(dev), 'item_id': pos[sl].to(dev)})
        la, lc = trainer.fused(inter)
        dist.all_reduce(trainer.optimizer.flat_grad)
        trainer.optimizer.flat_grad.mul_(1.0 / world)
        res.append((float(la), float(lc), trainer.optimizer.flat_grad.clone()))
        if vocab_parallel:
            model.eval()
            with torch.no_grad():
                out, _ = model._encode(inter['item_id_list'], inter['item_length'], need_attacked=False)
                v1, i1, r1 = A.ops.full_sort_topk(out, model.item_embedding.weight, k, inter['item_id'])
                v2, i2, r2 = trainer.vp.full_sort_topk(out, model.item_embedding.weight, k, inter['item_id'])
            scores = (out.double() @ model.item_embedding.weight.double().t()).float().cpu()
            ok, nbad = O.topk_equal_modulo_ties(i2.cpu(), i1.cpu(), scores)
            assert ok, nbad
            assert torch.equal(r1[:, -1], r2[:, -1])
    (a0, c0, g0), (a1, c1, g1) = res
    assert abs(a0 - a1) < 1e-5 * abs(a0) and abs(c0 - c1) < 1e-5 * abs(c0), (a0, a1, c0, c1)
    err = float((g0 - g1).abs().max()) / float(g0.abs().max())
    assert err < 2e-4, err
    dist.barrier()
    if rank == 0:
        print('multi-gpu check ok: world %d, grad rel err %.2e' % (world, err))
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == '__main__':
    main()
