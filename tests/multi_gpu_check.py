"""Run under torchrun with 2+ GPUs (scripts/gpu_multi.sh): the three multi-GPU modes of the trainer give the same step
  (a) batch data-parallel, replicated item table (dense all-reduce of the flat gradient)          -- the reference point
  (b) vocab-parallel logits / CE / top-k over a replicated table
  (c) vocab-parallel with SHARDED table storage (rows, gradient and Adam moments on the owner only; row exchange per step)
Checked: both losses, every gradient (the sharded table gradient gathered back to the reference layout), the parameters
after two optimisation steps through the CUDA-graph path, the sharded top-k against the local top-k, the gathered
checkpoint.  Prints 'multi-gpu check ok' on rank 0."""
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
    modes = [('dp', False, False), ('vocab-parallel', True, False), ('vocab-parallel + sharded table', True, True)]
    res = {}
    for name, vocab_parallel, shard in modes:
        config = make_config(A, cfg, cuda_graph=True, device=dev, checkpoint_dir='/tmp/acsr_mgpu_%d' % rank)
        model = A.ACSASRec(config, DS(V)).to(dev)
        model.load_state_dict({kk: v.to(dev) for kk, v in params.items()})
        model._debug_rand = noise
        trainer = A.ACSASRecTrainer(config, model)
        trainer.enable_data_parallel(vocab_parallel=vocab_parallel, shard_table=shard)
        model.train()
        inter = A.Interaction({'item_id_list': seq[sl].to(dev), 'item_length': ln[sl].to(dev), 'item_id': pos[sl].to(dev)})
        la, lc = trainer.fused(inter)
        trainer._allreduce_grads()
        grads = {}
        for n, p in model.named_parameters():
            gr = p.grad.detach()
            if n == 'item_embedding.weight' and shard:
                gr = trainer.vp.gather_full(gr)
            grads[n] = gr.clone()
        trainer.optimizer.zero_grad()
        # two optimisation steps through the trainer's own (CUDA-graph) path, NCCL collectives captured
        for _ in range(2):
            trainer.graphed_step(inter)
        sd = trainer._full_state_dict()
        entry = dict(la=float(la), lc=float(lc), grads=grads, sd=sd)
        model.eval()
        with torch.no_grad():
            val, idx, rec = model.full_sort_topk(inter, k, inter['item_id'])
            _, scores = model.full_sort_predict(inter)
        entry.update(idx=idx.cpu(), rec=rec.cpu(), scores=scores.cpu())
        if shard:
            assert tuple(model.item_embedding.weight.shape) == (trainer.vp.per, cfg['hidden_size'])
            trainer._save_checkpoint(0, verbose=False)
            if rank == 0:
                ck = torch.load(trainer.saved_model_file, map_location='cpu', weights_only=False)
                assert tuple(ck['state_dict']['item_embedding.weight'].shape) == (V, cfg['hidden_size'])
                st0 = ck['optimizer']['state']
                shapes = [tuple(st0[i]['exp_avg'].shape) for i in sorted(st0)]
                assert (V, cfg['hidden_size']) in shapes
        res[name] = entry
        del trainer, model
        torch.cuda.synchronize()
        dist.barrier()
    ref = res['dp']
    worst = 0.0
    for name in ('vocab-parallel', 'vocab-parallel + sharded table'):
        e = res[name]
        assert abs(e['la'] - ref['la']) < 1e-5 * abs(ref['la']) and abs(e['lc'] - ref['lc']) < 1e-5 * abs(ref['lc']), (name, e['la'], ref['la'])
        for n, gr in ref['grads'].items():
            scale = float(gr.abs().max())
            err = float((e['grads'][n] - gr).abs().max())
            assert err <= 2e-4 * scale + 1e-9, (name, n, err, scale)
            if scale > 1e-6:               # (a key-side bias has an exactly-zero true gradient: rounding noise on both sides)
                worst = max(worst, err / scale)
        for n, v in ref['sd'].items():
            # two Adam steps move every weight by ~2 lr: compare the updates
            upd_ref, upd = v - params[n], e['sd'][n] - params[n]
            big = ref['grads'][n].cpu().abs() > 1e-2 * ref['grads'][n].abs().max().cpu()      # Adam moves by ~lr * sign(g): skip noise-level gradients
            if bool(big.any()):
                assert float((upd_ref - upd)[big].abs().max()) < 1e-4, (name, n)
        ok, nbad = O.topk_equal_modulo_ties(e['idx'], ref['idx'], ref['scores'])
        assert ok, (name, nbad)
        assert torch.equal(e['rec'][:, -1], ref['rec'][:, -1])
        assert float((e['scores'] - ref['scores']).abs().max()) < 1e-4 * float(ref['scores'].abs().max()), name
    dist.barrier()
    # ---- sibling models: batch data-parallel through the same trainer; the captured autograd step contains the
    # NCCL all-reduce.  Every rank takes its own rows, so after three steps all ranks must hold the same parameters, and those must
    # equal a single-process run over the concatenated batch only up to the mean-of-means of the per-rank losses -- checked here:
    # rank-identical parameters, finite losses, parameters moved.
    sib = []
    for name in ('ACSSEPT', 'ACTiSASRec', 'AcBERT4Rec'):
        cfg2 = O.default_cfg(n_layers=2)
        cfg2.update(time_span=64, TIME_FIELD='timestamp', user_hidden_size=32, item_hidden_size=32, mask_ratio=0.2)
        config = make_config(A, cfg2, cuda_graph=True, device=dev, checkpoint_dir='/tmp/acsr_mgpu_%d' % rank, train_batch_size=B, eval_batch_size=B)
        config['model'] = name
        ds = A.data.SyntheticSequentialDataset(config, 3 * B, 300, seed=40 + rank)          # different rows on every rank
        torch.manual_seed(5)
        model = getattr(A, name)(config, ds).to(dev)
        trainer = getattr(A, name + 'Trainer')(config, model)
        trainer.enable_data_parallel()
        before = trainer.optimizer.flat_param.clone()
        model.train()
        la, lc = trainer._train_epoch(A.data.TrainDataLoader(config, ds, shuffle=False), 0)
        assert trainer._graph is not None and la == la and lc == lc
        flat = trainer.optimizer.flat_param
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        for r in range(1, world):
            assert torch.equal(gathered[0], gathered[r]), (name, 'rank %d diverged' % r)
        assert float((flat - before).abs().max()) > 0
        sib.append(name)
        del trainer, model
        torch.cuda.synchronize()
        dist.barrier()
    if rank == 0:
        print('multi-gpu check ok: world %d, 3 modes, worst grad rel err %.2e; data-parallel sibling models: %s' % (world, worst, ', '.join(sib)))
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == '__main__':
    main()
