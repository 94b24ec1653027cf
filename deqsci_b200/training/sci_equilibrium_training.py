"""Evaluation / training loops of the SCI path — drop-in for the reference's
training/sci_equilibrium_training.py (test_solver_sci :152-205, train_solver_sci :28-150).

test_solver_sci keeps the reference's protocol (drop/runner use only measurement 0; PSNR of the
clipped 8-frame cube; scene PSNR = mean over measurements; reported average = mean over scenes;
one [H,W,1]*255 array per frame in the returned dict) but reconstructs ALL measurements of a scene
in one batch: alpha is per sample in andersonexp and the sigma schedule resets once per scene either
way, so the per-measurement results are the same while the GPU sees B = #measurements."""
import os

import numpy as np
import torch

from ..distributed import allreduce_mean_gradients
from ..utils import cg_utils
from ..utils.metrics import peak_signal_noise_ratio


def tensor_to_np(tensor):
    return tensor.clip(0, 1).cpu().detach().unsqueeze(2).numpy() * 255.


def test_solver_sci(deep_eq_module, test_dataloader=None, save_img_path=None, verbose=True, save_image=True,
                    device=None):
    all_images = {}
    device = device or torch.device("cuda", torch.cuda.current_device())
    psnr_sum_for_avg, num_for_avg = 0, 0
    for ii, sample_batch in enumerate(test_dataloader):
        gt_batch = torch.as_tensor(sample_batch['gt']).to(device)
        y_batch = torch.as_tensor(sample_batch['meas']).to(device)
        Phi = torch.as_tensor(sample_batch['mask']).to(device)
        file_name = sample_batch['file']
        if isinstance(file_name, str):
            file_name = [file_name]
        if gt_batch.dim() == 3:      # dataset item without the DataLoader's batch dimension
            gt_batch, y_batch, Phi = gt_batch[None], y_batch[None], Phi[None]
        if ('drop' in file_name[0]) or ('runner' in file_name[0]):
            y_batch = y_batch[:, :, :, 0].unsqueeze(3)
        bsz, h, w, f = y_batch.shape
        T = Phi.shape[3]
        # [bsz,h,w,f] -> f*bsz independent measurements sharing the scene's mask
        y = y_batch.permute(3, 0, 1, 2).reshape(f * bsz, h, w).contiguous()
        Phi_b = Phi.repeat(f, 1, 1, 1)
        Phi_sum = cg_utils.Phi_sum_(Phi_b)
        with torch.no_grad():
            initial_point = cg_utils.initial_point(y, Phi_b, Phi_sum, None)
        reconstruction = deep_eq_module.forward(y, Phi_b, Phi_sum, initial_point=initial_point, train_flag=False)
        rec = reconstruction.clip(0, 1).cpu().detach().numpy().reshape(f, bsz, h, w, T)
        psnr_sum = 0
        for fi in range(f):
            gt = gt_batch[:, :, :, fi * T:(fi + 1) * T].cpu().numpy()
            psnr_sum += peak_signal_noise_ratio(gt, rec[fi])
            for frame_id in range(T):
                all_images[(save_img_path or '') + '%s_reconstruction_%d.png' % (file_name[0], fi * T + frame_id)] = \
                    rec[fi, 0, :, :, frame_id][:, :, None] * 255.
        current_psnr = psnr_sum / f
        psnr_sum_for_avg += current_psnr
        num_for_avg += 1
        if verbose:
            print(file_name, '  PSNR: %.2f dB' % current_psnr)
    avg_psnr = psnr_sum_for_avg / num_for_avg
    if verbose:
        print('---------------------------------', 'Total Average PSNR: %.2f dB' % avg_psnr)
    if save_image and save_img_path is not None:
        import cv2
        os.makedirs(save_img_path, exist_ok=True)
        for k in all_images:
            cv2.imwrite(k, all_images[k])
    return avg_psnr, all_images


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
    for epoch in range(start_epoch, n_epochs):
        for ii, sample_batch in enumerate(train_dataloader):
            optimizer.zero_grad()
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
