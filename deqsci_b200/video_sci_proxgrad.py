"""Entry script — same flags as the reference's video_sci_proxgrad.py:23-49 (including the
string-truthy --inference and the untyped --and_maxiters), running the B200 path.

    python -m deqsci_b200.video_sci_proxgrad --savepath ./save/test_ffdnet/ --testpath ./data/test_gray/ \\
        --loadpath ./models/ffdnet.ckpt --denoiser ffdnet --and_maxiters 180 --inference True

Denoisers on the path built here: ffdnet, SimpleCNN, RealSN_SimpleCNN (the three shipped
checkpoints), DnCNN and RealSN_DnCNN (the 17-layer BatchNorm DnCNNs, same conv stack).  The U-Net / ResNet
choices of the reference are outside the DE-GAP scope and raise NotImplementedError."""
import argparse
import os

import torch
import torch.optim as optim


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpu_ids', default='0')
    parser.add_argument('--n_epochs', default=80)
    parser.add_argument('--batch_size', type=int, default=1)
    parser.add_argument('--and_maxiters', default=100)
    parser.add_argument('--and_beta', type=float, default=1.0)
    parser.add_argument('--and_m', type=int, default=5)
    parser.add_argument('--lr', type=float, default=0.0001)
    parser.add_argument('--etainit', type=float, default=0.9)
    parser.add_argument('--lr_gamma', type=float, default=0.9)
    parser.add_argument('--sched_step', type=int, default=10)
    parser.add_argument('--savepath', default="./save/test/")
    parser.add_argument('--trainpath', default="./data/train/")
    parser.add_argument('--testpath', default="./data/test_gray/")
    parser.add_argument('--loadpath', default='')
    parser.add_argument('--denoiser', default='ffdnet')
    parser.add_argument('--inference', default=False, help='turning model to training or testing mode.')
    parser.add_argument('--print_every_n_steps', type=int, default=1)
    parser.add_argument('--save_every_n_steps', type=int, default=50)
    parser.add_argument('--sigma', type=int, default=0)
    return parser


def build_denoiser(name):
    if name == 'ffdnet':
        from .networks.ffdnet.models import FFDNet
        return FFDNet(num_input_channels=1, tag='ffdnet')
    if name == 'SimpleCNN':
        from .networks.provable.model.SimpleCNN_models import DnCNN
        return DnCNN(1, num_of_layers=4, lip=0.0, no_bn=True, tag='denoiser')
    if name == 'RealSN_SimpleCNN':
        from .networks.provable.model.SimpleCNN_models import DnCNN
        return DnCNN(1, num_of_layers=4, lip=1.0, no_bn=True, tag='denoiser')
    if name == 'DnCNN':
        from .networks.provable.model.models import DnCNN
        return DnCNN(channels=1, num_of_layers=17, tag='denoiser')
    if name == 'RealSN_DnCNN':
        from .networks.provable.model.realSN_models import DnCNN
        return DnCNN(channels=1, num_of_layers=17, tag='denoiser')
    raise NotImplementedError('unknown denoiser! (%r is not on the DE-GAP path of this build)' % name)


def main(argv=None):
    args = build_parser().parse_args(argv)
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", args.gpu_ids)
    from .solvers import new_equilibrium_utils_yaping as eq_utils
    from .solvers.equilibrium_solvers_yaping import EquilibriumProxGradSCI
    from .training import sci_equilibrium_training
    from .utils.cg_utils import A_torch_, At_torch_
    from .utils.sci_dataloader import SCITestDataset, SCITrainingDatasetSubset

    inference = args.inference                      # any non-empty string is truthy, as in the reference
    max_iters = int(args.and_maxiters)
    save_location = args.savepath
    save_model_path = save_location + 'model/'
    save_test_img_path = save_location + 'img/test/'
    for path in (save_model_path, save_location + 'img/train/', save_test_img_path):
        os.makedirs(path, exist_ok=True)
    print('cuda', torch.cuda.is_available())

    test_dataloader = torch.utils.data.DataLoader(SCITestDataset(args.testpath), batch_size=1, shuffle=False,
                                                  drop_last=True)
    learned_component = build_denoiser(args.denoiser)
    if inference:
        learned_component.eval()
    solver = EquilibriumProxGradSCI(A=A_torch_, At=At_torch_, nonlinear_operator=learned_component, eta=0.2,
                                    minval=-1, maxval=1).cuda()
    optimizer = optim.Adam(params=solver.parameters(), lr=float(args.lr))
    scheduler = optim.lr_scheduler.StepLR(optimizer=optimizer, step_size=int(args.sched_step),
                                          gamma=float(args.lr_gamma))
    load_location = args.loadpath
    if args.sigma:
        load_location = "./networks/provable/Pretrained_models/" + args.denoiser + "_noise" + str(args.sigma) + ".pth"
    start_epoch = 0
    if os.path.exists(load_location):
        saved_dict = torch.load(load_location, map_location='cuda')
        start_epoch = saved_dict['epoch'] + 1
        sd = {(k[7:] if k.startswith('module.') else k): v for (k, v) in saved_dict['solver_state_dict'].items()}
        solver.load_state_dict(sd)
        print('loaded dict!')
    lossfunction = torch.nn.MSELoss(reduction='mean')
    deep_eq_module = eq_utils.DEQFixedPoint(solver, eq_utils.andersonexp, m=int(args.and_m),
                                            beta=float(args.and_beta), lam=1e-2, max_iter=max_iters, tol=1e-5)
    if not inference:
        dataset = SCITrainingDatasetSubset(args.trainpath + 'gt/', args.trainpath + 'measurement/',
                                           args.trainpath + 'mask.mat')
        dataloader = torch.utils.data.DataLoader(dataset=dataset, batch_size=int(args.batch_size), shuffle=True,
                                                 drop_last=True, pin_memory=True)
        sci_equilibrium_training.train_solver_sci(
            single_iterate_solver=solver, train_dataloader=dataloader, test_dataloader=test_dataloader,
            optimizer=optimizer, save_model_path=save_model_path, deep_eq_module=deep_eq_module,
            loss_function=lossfunction, n_epochs=int(args.n_epochs), scheduler=scheduler,
            print_every_n_steps=args.print_every_n_steps, save_every_n_steps=args.save_every_n_steps,
            start_epoch=start_epoch, test_img_path=save_test_img_path)
    else:
        cur_psnr, all_images = sci_equilibrium_training.test_solver_sci(
            test_dataloader=test_dataloader, deep_eq_module=deep_eq_module, save_img_path=save_test_img_path)
        return cur_psnr


if __name__ == "__main__":
    main()
