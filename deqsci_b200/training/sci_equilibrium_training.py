"""Evaluation / training loops of the SCI path — drop-in for the reference's
training/sci_equilibrium_training.py (test_solver_sci :152-205, train_solver_sci :28-150).

test_solver_sci keeps the reference's protocol (drop/runner use only measurement 0; PSNR of the clipped
8-frame cube; scene PSNR = mean over its measurements; reported average = mean over scenes; one [H,W,1]*255
array per frame in the returned dict) but packs ALL measurements of ALL scenes of one shape into ONE batched
solve: alpha is per sample in andersonexp and every measurement starts its own sigma schedule at
call 0, so each sample's trajectory is the one the reference's batch-1 loop computes -- as long as the stopping
test never fires.  The reference tests the residual of the measurement it is solving (batch 1); the batched solve
tests the whole-batch residual and also reports the smallest per-sample residual it saw
(DEQFixedPoint.forward_min_sample_res): if either is below `tol`, some measurement would have stopped on its
own, and the group is re-solved measurement by measurement exactly as the reference does.  PSNR (and SSIM, kept
in `test_solver_sci.last_metrics`) are reduced on the device; PNGs are written by a background thread."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from ..distributed import GradientSynchronizer, allreduce_mean_gradients
from ..utils import cg_utils
from ..utils.metrics import peak_signal_noise_ratio, ssim


def tensor_to_np(tensor):
    return tensor.clip(0, 1).cpu().detach().unsqueeze(2).numpy() * 255.


def _device_psnr(rec, gt):
    """PSNR of every [H,W,T] cube of the batch on the device (skimage float rule with data_range 1: the reference
    passes the CLIPPED reconstruction as image_true, :182-183).  fp64 accumulation, one [B] result."""
    mse = ((rec.clip(0, 1).double() - gt.double()) ** 2).mean(dim=(1, 2, 3))
    return 10.0 * torch.log10(1.0 / mse)


def test_solver_sci(deep_eq_module, test_dataloader=None, save_img_path=None, verbose=True, save_image=True,
                    device=None):
    all_images = {}
    device = device or torch.device("cuda", torch.cuda.current_device())
    # ---- gather every scene (the loader yields one scene per item), group by cube shape ---------------------
    scenes, groups = [], {}
    for ii, sample_batch in enumerate(test_dataloader):
        gt_batch = torch.as_tensor(sample_batch['gt']).to(device, non_blocking=True)
        y_batch = torch.as_tensor(sample_batch['meas']).to(device, non_blocking=True)
        Phi = torch.as_tensor(sample_batch['mask']).to(device, non_blocking=True)
        file_name = sample_batch['file']
        if isinstance(file_name, str):
            file_name = [file_name]
        if gt_batch.dim() == 3:      # dataset item without the DataLoader's batch dimension
            gt_batch, y_batch, Phi = gt_batch[None], y_batch[None], Phi[None]
        if ('drop' in file_name[0]) or ('runner' in file_name[0]):
            y_batch = y_batch[:, :, :, 0].unsqueeze(3)
        bsz, h, w, f = y_batch.shape
        T = Phi.shape[3]
        # [bsz,h,w,f] -> f*bsz independent measurements sharing the scene's mask; gt cube of measurement fi
        y = y_batch.permute(3, 0, 1, 2).reshape(f * bsz, h, w)
        gt = torch.cat([gt_batch[:, :, :, fi * T:(fi + 1) * T] for fi in range(f)], 0)
        sc = {"name": file_name[0], "f": f, "bsz": bsz, "y": y, "Phi": Phi.repeat(f, 1, 1, 1), "gt": gt}
        scenes.append(sc)
        groups.setdefault((h, w, T), []).append(sc)
    tol = float(deep_eq_module.kwargs.get("tol", 1e-5)) if hasattr(deep_eq_module, "kwargs") else 0.0
    writer = ThreadPoolExecutor(max_workers=4) if (save_image and save_img_path is not None) else None
    pending = []
    if writer is not None:
        os.makedirs(save_img_path, exist_ok=True)
    metrics = {}
    for (h, w, T), group in groups.items():
        y = torch.cat([sc["y"] for sc in group], 0).contiguous()
        Phi_b = torch.cat([sc["Phi"] for sc in group], 0).contiguous()
        gt = torch.cat([sc["gt"] for sc in group], 0)
        with torch.no_grad():
            Phi_sum = cg_utils.Phi_sum_(Phi_b)
            initial_point = cg_utils.initial_point(y, Phi_b, Phi_sum, None)
            reconstruction = deep_eq_module.forward(y, Phi_b, Phi_sum, initial_point=initial_point, train_flag=False)
            res_b = getattr(deep_eq_module, "forward_res", None)
            res_s = getattr(deep_eq_module, "forward_min_sample_res", None)
            fired = [r for r in (res_b, res_s) if r is not None and r < tol]
            if fired and y.shape[0] > 1:
                # some measurement alone would have met the stopping test: the reference's per-measurement loop
                parts = []
                for i in range(y.shape[0]):
                    yi, Pi, Si = y[i:i + 1].contiguous(), Phi_b[i:i + 1].contiguous(), Phi_sum[i:i + 1].contiguous()
                    parts.append(deep_eq_module.forward(yi, Pi, Si, initial_point=cg_utils.initial_point(yi, Pi, Si, None),
                                                        train_flag=False))
                reconstruction = torch.cat(parts, 0)
            psnr = _device_psnr(reconstruction, gt)
            rc = reconstruction.clip(0, 1)
            ss = ssim(rc.permute(0, 3, 1, 2).contiguous(), gt.permute(0, 3, 1, 2).contiguous(), size_average=False)
            psnr_h, ssim_h = psnr.cpu().tolist(), ss.cpu().tolist()             # one small D2H per group
            rec = rc.cpu().numpy()
        off = 0
        for sc in group:
            n = sc["f"] * sc["bsz"]
            sc["psnr"] = float(np.mean(psnr_h[off:off + n]))
            metrics[sc["name"]] = {"psnr": psnr_h[off:off + n], "ssim": ssim_h[off:off + n]}
            for fi in range(sc["f"]):
                for frame_id in range(T):
                    key = (save_img_path or '') + '%s_reconstruction_%d.png' % (sc["name"], fi * T + frame_id)
                    img = rec[off + fi * sc["bsz"], :, :, frame_id][:, :, None] * 255.
                    all_images[key] = img
                    if writer is not None:
                        pending.append(writer.submit(_imwrite, key, img))
            off += n
    psnr_sum_for_avg, num_for_avg = 0, 0
    for sc in scenes:                          # report in loader order, as the reference prints
        psnr_sum_for_avg += sc["psnr"]
        num_for_avg += 1
        if verbose:
            print([sc["name"]], '  PSNR: %.2f dB' % sc["psnr"])
    avg_psnr = psnr_sum_for_avg / num_for_avg
    if verbose:
        print('---------------------------------', 'Total Average PSNR: %.2f dB' % avg_psnr)
    if writer is not None:
        for fut in pending:
            fut.result()
        writer.shutdown()
    test_solver_sci.last_metrics = metrics
    return avg_psnr, all_images


test_solver_sci.last_metrics = {}


def _imwrite(path, img):
    import cv2
    cv2.imwrite(path, img)


def train_solver_sci(single_iterate_solver, train_dataloader, test_dataloader, optimizer, save_model_path,
                     deep_eq_module, loss_function, n_epochs, use_dataparallel=False, scheduler=None,
                     print_every_n_steps=1, save_every_n_steps=50, start_epoch=0, train_img_path=None,
                     test_img_path=None, best_img_path=None, tflog_path=None, device=None):
    """Implicit-differentiation training loop (reference :28-150): forward solve, MSE loss,
    loss.backward() through DEQFixedPoint's hook, Adam step, periodic test + best/epoch checkpoints
    with the reference's dict keys.  When torch.distributed is initialised the parameter gradients
    are averaged over ranks (one flat NCCL all-reduce) between backward() and step()."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    best_psnr = 0.0
    # a plain Adam (the reference's optimizer) is taken over by the fused exchange-and-update kernel; anything
    # else keeps the generic flat all-reduce + optimizer.step()
    sync = GradientSynchronizer.adopt(optimizer)
    for epoch in range(start_epoch, n_epochs):
        for ii, sample_batch in enumerate(train_dataloader):
            sync.zero_grad() if sync is not None else optimizer.zero_grad()
            gt = sample_batch['gt'].to(device)
            y = sample_batch['meas'].to(device)
            Phi = sample_batch['mask'].to(device)
            Phi_sum = torch.sum(Phi, dim=3)
            Phi_sum[Phi_sum == 0] = 1
            with torch.no_grad():
                x0 = cg_utils.initial_point(y, Phi, Phi_sum, gt)
            reconstruction = deep_eq_module.forward(y, Phi, Phi_sum, initial_point=x0)
            loss = loss_function(reconstruction, gt)
            if torch.isnan(loss):
                continue
            loss.backward()
            if sync is not None:
                sync.step()                    # all-reduce(mean) + Adam: one kernel (csrc/optim.cu)
            else:
                allreduce_mean_gradients(single_iterate_solver.parameters())
                optimizer.step()
            if ii % print_every_n_steps == 0:
                psnr = peak_signal_noise_ratio(gt.cpu().numpy(), reconstruction.clip(0, 1).cpu().detach().numpy())
                print("Epoch %d, step %d: loss %.6f PSNR %.2f dB" % (epoch, ii, float(loss), psnr), flush=True)
            if test_dataloader is not None and ii % save_every_n_steps == 0 and ii > 0:
                was_training = single_iterate_solver.nonlinear_op.training
                single_iterate_solver.nonlinear_op.eval()
                cur, _ = test_solver_sci(deep_eq_module, test_dataloader, test_img_path, verbose=False,
                                         save_image=False, device=device)
                single_iterate_solver.nonlinear_op.train(was_training)
                if cur > best_psnr and save_model_path:
                    best_psnr = cur
                    _save(single_iterate_solver, optimizer, scheduler, epoch, os.path.join(save_model_path, 'best.ckpt'))
        if scheduler is not None:
            scheduler.step()
        if save_model_path:
            _save(single_iterate_solver, optimizer, scheduler, epoch,
                  os.path.join(save_model_path, 'epoch_%d.ckpt' % epoch))


def _save(solver, optimizer, scheduler, epoch, path):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save({'solver_state_dict': solver.state_dict(), 'epoch': epoch,
                'optimizer_state_dict': optimizer.state_dict(),
                'scheduler_state_dict': scheduler.state_dict() if scheduler is not None else None}, path)
