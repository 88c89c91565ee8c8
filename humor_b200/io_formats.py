"""On-disk formats either side of the fitting path, byte-compatible with what the
reference reads and writes so that its viz_fitting_rgb.py / eval_fitting_*.py consume the outputs unchanged.
Plain host code (numpy + json): nothing here touches the GPU.

  in   OpenPose ``*_keypoints.json`` -> joints2d (T,25,3)      fitting_utils.py:21-37 (read_keypoints),
       split of a video into overlapping sub-sequences          datasets/rgb_dataset.py:63-158
       ``prior_gmm.npz`` {weights, means, covariances}          run_fitting.py:251-258
       checkpoint dict {'model', 'optim', 'epoch', ...}         utils/torch.py:36-82 (load_state)
  out  per sub-sequence ``stage3_results.npz`` (+ ``_prior``), ``gt_results.npz`` / ``proxd_results.npz``,
       ``observations.npz``, ``meta.txt``                       fitting_utils.py:274-395, run_fitting.py:373-382
       stitched ``final_results/``                              fitting_utils.py:398-523
"""
import glob
import json
import math
import os
import shutil

import numpy as np

OP_NUM_JOINTS = 25
DEFAULT_GROUND = [0.0, -1.0, 0.0, -0.5]                      # datasets/rgb_dataset.py:16
DEFAULT_FOCAL_LEN = (1060.531764702488, 1060.3856705041237)  # fitting_utils.py:19


# ------------------------------------------------------------------------------------------------ inputs
def read_keypoints(keypoint_fn):
    """Body keypoints of the FIRST person of an OpenPose json -> (25,3) float64 (x, y, confidence); zeros when the
    frame has no detection (fitting_utils.py:21-37; the reference's ``np.float`` is float64)."""
    with open(keypoint_fn) as f:
        data = json.load(f)
    if len(data['people']) == 0:
        print('WARNING: Found no keypoints in %s! Returning zeros!' % (keypoint_fn))
        return np.zeros((OP_NUM_JOINTS, 3), dtype=np.float64)
    return np.array(data['people'][0]['pose_keypoints_2d'], dtype=np.float64).reshape([-1, 3])


def split_intervals(num_frames, seq_len=None, overlap_len=None):
    """Sub-sequence frame intervals of a video (rgb_dataset.py:74-99): ``ceil((F-o)/(L-o))`` windows of length L; the
    overlap is first grown evenly so that the windows cover the video with as little excess as possible, and the
    remaining r excess frames are absorbed by one extra frame of overlap on the first r windows.
    Returns (list of (start, end), overlap actually used)."""
    if seq_len is None or overlap_len is None:
        return [(0, num_frames)], overlap_len
    num_seqs = math.ceil((num_frames - overlap_len) / (seq_len - overlap_len))
    r = seq_len * num_seqs - overlap_len * (num_seqs - 1) - num_frames
    if num_seqs > 1:
        overlap_len = overlap_len + r // (num_seqs - 1)
    # num_seqs == 1: the reference divides by zero here (a video no longer than one window must be run unsplit)
    r = seq_len * num_seqs - overlap_len * (num_seqs - 1) - num_frames
    out = []
    s = 0
    for i in range(num_seqs):
        out.append((s, s + seq_len))
        s += seq_len - (overlap_len + (1 if i < r else 0))
    return out, overlap_len


def keypoint_paths(joints2d_path):
    return sorted(glob.glob(os.path.join(joints2d_path, '*_keypoints.json')))


def load_rgb_video(joints2d_path, cam_mat, seq_len=None, overlap_len=None, floor_plane=None, video_name='rgb_video'):
    """What RGBVideoDataset.load_data/__getitem__ hand to run_fitting.py for an OpenPose directory, stacked over the
    sub-sequences: observed {'joints2d' (B,T,25,3) f32, 'floor_plane' (B,4) f64, 'seq_interval' (B,2) i32} and
    gt {'cam_matx' (B,3,3) f32, 'name' [B]}."""
    paths = keypoint_paths(joints2d_path)
    intervals, _ = split_intervals(len(paths), seq_len, overlap_len)
    fp = np.array(DEFAULT_GROUND if floor_plane is None else floor_plane, dtype=np.float64)
    j2d = [np.stack([read_keypoints(f) for f in paths[s:e]], 0) for s, e in intervals]
    obs = {'joints2d': np.stack(j2d, 0).astype(np.float32), 'floor_plane': np.stack([fp] * len(intervals), 0),
           'seq_interval': np.asarray(intervals, np.int32)}
    gt = {'cam_matx': np.stack([np.asarray(cam_mat, np.float32)] * len(intervals), 0),
          'name': [video_name + '_' + '%04d' % i for i in range(len(intervals))]}
    return obs, gt


def load_gmm(init_motion_prior_dir):
    """``prior_gmm.npz`` -> (weights (K,), means (K,D), covariances (K,D,D)) float32 (run_fitting.py:251-258)."""
    path = os.path.join(init_motion_prior_dir, 'prior_gmm.npz')
    if not os.path.exists(path):
        raise FileNotFoundError('Could not find init motion state prior at given directory! (%s)' % path)
    r = np.load(path)
    return tuple(np.asarray(r[k], np.float32) for k in ('weights', 'means', 'covariances'))


def load_state(load_path, model, map_location=None, ignore_keys=None):
    """utils/torch.py:44-82: checkpoint dict with key 'model' (optionally trained under DataParallel: 'module.' prefix),
    loaded non-strictly; returns (epoch, min_val_loss, min_train_loss)."""
    import torch
    ckpt = torch.load(load_path, map_location=map_location, weights_only=False)
    sd = ckpt['model']
    if len(sd) and next(iter(sd)).split('.')[0] == 'module':
        sd = {'.'.join(k.split('.')[1:]): v for k, v in sd.items() if k.split('.')[0] == 'module'}
    if ignore_keys is not None:
        sd = {k: v for k, v in sd.items() if k.split('.')[0] not in ignore_keys}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    if ignore_keys is not None:
        missing = [k for k in missing if k.split('.')[0] not in ignore_keys]
        unexpected = [k for k in unexpected if k.split('.')[0] not in ignore_keys]
    if missing:
        print('WARNING: The following keys could not be found in the given state dict - ignoring...\n%s' % missing)
    if unexpected:
        print('WARNING: The following keys were found in the given state dict but not in the current model - ignoring...\n%s'
              % unexpected)
    return ckpt.get('epoch'), ckpt.get('min_val_loss'), ckpt.get('min_train_loss', float('Inf'))


# ------------------------------------------------------------------------------------------------ outputs
def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, 'detach') else np.asarray(x)


def write_meta(res_out_paths, body_model_path, gt_body_paths=None):
    """meta.txt (run_fitting.py:373-382)."""
    for b, p in enumerate(res_out_paths):
        with open(os.path.join(p, 'meta.txt'), 'w') as f:
            f.write('optim_bm %s\n' % body_model_path)
            f.write('gt_bm %s\n' % (body_model_path if gt_body_paths is None else gt_body_paths[b]))


def save_optim_result(cur_res_out_paths, optim_result, per_stage_results, gt_data, observed_data, data_type,
                      optim_floor=True, obs_img_paths=None, obs_mask_paths=None):
    """fitting_utils.py:274-395, same files and keys."""
    res = {k: _np(optim_result[k]) for k in ('betas', 'trans', 'root_orient', 'pose_body')}
    contacts = _np(optim_result['contacts']) if 'contacts' in optim_result else None
    floor = _np(optim_result['floor_plane']) if 'floor_plane' in optim_result else None
    for b, p in enumerate(cur_res_out_paths):
        d = {k: v[b] for k, v in res.items()}
        if contacts is not None:
            d['contacts'] = contacts[b]
        if floor is not None:
            d['floor_plane'] = floor[b]
        np.savez(os.path.join(p, 'stage3_results.npz'), **d)
    if 'stage3' in per_stage_results and optim_floor:
        ptr, pro = _np(per_stage_results['stage3']['prior_trans']), _np(per_stage_results['stage3']['prior_root_orient'])
        for b, p in enumerate(cur_res_out_paths):
            d = {'betas': res['betas'][b], 'trans': ptr[b], 'root_orient': pro[b], 'pose_body': res['pose_body'][b]}
            if contacts is not None:
                d['contacts'] = contacts[b]
            np.savez(os.path.join(p, 'stage3_results_prior.npz'), **d)
    prox = data_type in ['PROX-RGB', 'PROX-RGBD']
    if all(k in gt_data for k in ('betas', 'trans', 'root_orient', 'pose_body')):
        gb = _np(gt_data['betas'])
        if not prox:
            gb = gb[:, 0]
        gt = {'trans': _np(gt_data['trans']), 'root_orient': _np(gt_data['root_orient']), 'pose_body': _np(gt_data['pose_body'])}
        gc = _np(gt_data['contacts']) if 'contacts' in gt_data else None
        cam = _np(gt_data['cam_matx']) if 'cam_matx' in gt_data else None
        for b, p in enumerate(cur_res_out_paths):
            d = {'betas': gb[b], 'trans': gt['trans'][b], 'root_orient': gt['root_orient'][b], 'pose_body': gt['pose_body'][b]}
            if gc is not None:
                d['contacts'] = gc[b]
            if cam is not None:
                d['cam_mtx'] = cam[b]
            np.savez(os.path.join(p, 'proxd_results.npz' if prox else 'gt_results.npz'), **d)
            if prox:
                np.savez(os.path.join(p, 'gt_results.npz'), cam_mtx=cam[b])
    elif 'joints3d' in gt_data:
        gj = _np(gt_data['joints3d'])
        cam = _np(gt_data['cam_matx']) if 'cam_matx' in gt_data else None
        occ = _np(gt_data['occlusions']) if 'occlusions' in gt_data else None
        for b, p in enumerate(cur_res_out_paths):
            d = {'joints3d': gj[b]}
            if cam is not None:
                d['cam_mtx'] = cam[b]
            if occ is not None:
                d['occlusions'] = occ[b]
            np.savez(os.path.join(p, 'gt_results.npz'), **d)
    elif 'cam_matx' in gt_data:
        cam = _np(gt_data['cam_matx'])
        for b, p in enumerate(cur_res_out_paths):
            np.savez(os.path.join(p, 'gt_results.npz'), cam_mtx=cam[b])
    obs = {k: _np(v) for k, v in observed_data.items() if k != 'prev_batch_overlap_res'}
    for b, p in enumerate(cur_res_out_paths):
        d = {k: v[b] for k, v in obs.items() if k not in ['RGB']}
        if obs_img_paths is not None:
            d['img_paths'] = [t[b] for t in obs_img_paths]
        if obs_mask_paths is not None:
            d['mask_paths'] = [t[b] for t in obs_mask_paths]
        np.savez(os.path.join(p, 'observations.npz'), **d)


def load_res(result_dir, file_name):
    """fitting_utils.py:526-535."""
    path = os.path.join(result_dir, file_name)
    if not os.path.exists(path):
        return None
    r = np.load(path, allow_pickle=True)
    return {k: r[k] for k in r.files}


def stitch_subsequences(seq_intervals, all_res_dirs):
    """The concatenation half of save_rgb_stitched_result (fitting_utils.py:401-446): every sub-sequence after the first
    drops the frames it shares with its predecessor.  Returns dict with betas (F,16), trans/root_orient/pose_body (F,.),
    contacts (F,22), floor_planes (S,.), joints2d (F,25,3), img_paths [F] (None when absent), cam_mtx."""
    overlaps = [0] + [seq_intervals[i][1] - seq_intervals[i + 1][0] for i in range(len(seq_intervals) - 1)]
    cat = {k: [] for k in ('betas', 'trans', 'root_orient', 'pose_body', 'contacts', 'joints2d')}
    floors, img_paths, cam = [], None, None
    for i, d in enumerate(all_res_dirs):
        if i >= len(overlaps):          # an extra directory from padding the batch to an even size
            break
        r = load_res(d, 'stage3_results.npz')
        T = r['trans'].shape[0]
        o = overlaps[i]
        betas = r['betas']
        if betas.ndim == 1:
            betas = np.broadcast_to(betas[None], (T, betas.shape[0]))        # prep_res, fitting_utils.py:537-550
        cat['betas'].append(betas[o:])
        for k in ('trans', 'root_orient', 'pose_body', 'contacts'):
            cat[k].append(r[k][o:])
        floors.append(np.asarray(r['floor_plane']).reshape(1, -1))
        if cam is None:
            g = load_res(d, 'gt_results.npz')
            cam = None if g is None else g.get('cam_mtx')
        ob = load_res(d, 'observations.npz')
        cat['joints2d'].append(ob['joints2d'][o:])
        if 'img_paths' in ob:
            img_paths = (img_paths or []) + list(ob['img_paths'][o:])
    out = {k: np.concatenate(v, 0) for k, v in cat.items()}
    out['floor_planes'] = np.concatenate(floors, 0)
    out['img_paths'], out['cam_mtx'] = img_paths, cam
    return out


def save_rgb_stitched_result(seq_intervals, all_res_out_paths, res_out_path, prior_frame=None):
    """final_results/ of fitting_utils.py:398-523: meta.txt, gt_results.npz, observations.npz, stage3_results.npz and —
    when ``prior_frame`` (a callable: stitched dict -> (prior_trans (F,3), prior_root_orient (F,3))) is given —
    stage3_results_prior.npz.  The reference derives that prior-frame copy by running SMPL on the stitched sequence and
    applying the cam->prior transform of frame 0 (fitting_utils.py:482-523); MotionOptimizer.apply_cam2prior is the
    callable to pass when a GPU is at hand, the file formats themselves need none."""
    final = os.path.join(res_out_path, 'final_results')
    os.makedirs(final, exist_ok=True)
    st = stitch_subsequences(seq_intervals, all_res_out_paths)
    shutil.copyfile(os.path.join(all_res_out_paths[0], 'meta.txt'), os.path.join(final, 'meta.txt'))
    np.savez(os.path.join(final, 'gt_results.npz'), cam_mtx=st['cam_mtx'])
    np.savez(os.path.join(final, 'observations.npz'), joints2d=st['joints2d'], img_paths=st['img_paths'])
    np.savez(os.path.join(final, 'stage3_results.npz'), betas=st['betas'], trans=st['trans'], root_orient=st['root_orient'],
             pose_body=st['pose_body'], floor_plane=st['floor_planes'][0], contacts=st['contacts'])
    if prior_frame is not None:
        ptr, pro = prior_frame(st)
        np.savez(os.path.join(final, 'stage3_results_prior.npz'), betas=st['betas'], trans=_np(ptr), root_orient=_np(pro),
                 pose_body=st['pose_body'], contacts=st['contacts'])
    return final
