"""ORACLE (test infrastructure) — restatement of smplx==0.1.28 SMPL+H forward.

smplx is a third-party dependency of the reference (requirements.txt:10), absent from
/root/reference and not installable offline.  This restates its published algorithm
(``smplx.lbs.lbs`` + ``SMPLH.forward`` + ``VertexJointSelector``) as SURVEY.md Appendix A.1
records it, anchored on the reference's call sites:
  * constructor arguments      humor/body_model/body_model.py:37-68
  * forward call + fields read humor/body_model/body_model.py:78-110
  * batch_rodrigues            humor/utils/transforms.py:139-170 (verbatim twin of smplx's)
Works in the dtype of its inputs (fp32 for parity runs, fp64 for tight checks).
"""
import numpy as np
import torch

# smplx.vertex_ids.vertex_ids['smplh'] in VertexJointSelector order
EXTRA_VERTEX_IDS = [332, 6260, 2800, 4071, 583,             # nose reye leye rear lear
                    3216, 3226, 3387, 6617, 6624, 6787,     # LBigToe LSmallToe LHeel RBigToe RSmallToe RHeel
                    2746, 2319, 2445, 2556, 2673,           # l thumb index middle ring pinky
                    6191, 5782, 5905, 6016, 6133]           # r thumb index middle ring pinky


def rodrigues(aa):
    """(M,3) axis-angle -> (M,3,3). transforms.py:139-170: angle = ||r + 1e-8||."""
    angle = torch.norm(aa + 1e-8, dim=1, keepdim=True)
    d = aa / angle
    c, s = torch.cos(angle)[:, :, None], torch.sin(angle)[:, :, None]
    z = torch.zeros_like(d[:, 0])
    K = torch.stack([z, -d[:, 2], d[:, 1], d[:, 2], z, -d[:, 0], -d[:, 1], d[:, 0], z], 1).view(-1, 3, 3)
    eye = torch.eye(3, dtype=aa.dtype, device=aa.device)[None]
    return eye + s * K + (1 - c) * torch.bmm(K, K)


class SMPLHOracle:
    """Holds the model arrays the way smplx does after BodyModel.__init__ preprocessed them."""

    def __init__(self, asset, num_betas=16, dtype=torch.float32, device='cpu'):
        t = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype, device=device)
        self.v_template = t(asset['v_template'])                          # (V,3)
        self.shapedirs = t(np.asarray(asset['shapedirs'])[:, :, :num_betas])  # (V,3,nb)
        pd = np.asarray(asset['posedirs'])
        self.posedirs = t(pd.reshape(-1, pd.shape[-1]).T)                 # (459, V*3)
        self.J_regressor = t(asset['J_regressor'])                        # (52,V)
        self.lbs_weights = t(asset['weights'])                            # (V,52)
        par = np.asarray(asset['kintree_table'])[0].astype(np.int64).copy()
        par[0] = -1
        self.parents = par
        self.faces = torch.as_tensor(np.asarray(asset['f']).astype(np.int64), device=device)
        self.dtype, self.device = dtype, device

    def lbs(self, betas, full_pose):
        """smplx.lbs.lbs: returns (verts (N,V,3), posed joints (N,52,3)) before translation."""
        N = betas.shape[0]
        v_shaped = self.v_template[None] + torch.einsum('bl,mkl->bmk', betas, self.shapedirs)
        J = torch.einsum('bik,ji->bjk', v_shaped, self.J_regressor)
        R = rodrigues(full_pose.reshape(-1, 3)).view(N, -1, 3, 3)
        eye = torch.eye(3, dtype=R.dtype, device=R.device)
        pose_feat = (R[:, 1:] - eye).reshape(N, -1)
        v_posed = v_shaped + (pose_feat @ self.posedirs).view(N, -1, 3)
        # batch_rigid_transform
        par = self.parents
        rel = J.clone()
        rel[:, 1:] = rel[:, 1:] - J[:, par[1:]]
        G = torch.zeros(N, J.shape[1], 4, 4, dtype=R.dtype, device=R.device)
        G[:, :, :3, :3] = R
        G[:, :, :3, 3] = rel
        G[:, :, 3, 3] = 1.0
        chain = [G[:, 0]]
        for i in range(1, J.shape[1]):
            chain.append(chain[par[i]] @ G[:, i])
        Gw = torch.stack(chain, 1)
        J_posed = Gw[:, :, :3, 3]
        Jh = torch.cat([J, torch.zeros_like(J[..., :1])], -1)[..., None]      # (N,52,4,1)
        corr = torch.matmul(Gw, Jh)                                           # (N,52,4,1)
        A = Gw.clone()
        A[:, :, :, 3:] = A[:, :, :, 3:] - corr
        T = (self.lbs_weights[None].expand(N, -1, -1) @ A.view(N, -1, 16)).view(N, -1, 4, 4)
        vh = torch.cat([v_posed, torch.ones_like(v_posed[..., :1])], -1)[..., None]
        verts = torch.matmul(T, vh)[:, :, :3, 0]
        return verts, J_posed

    def forward(self, betas, global_orient, body_pose, transl,
                left_hand_pose=None, right_hand_pose=None):
        """SMPLH.forward(use_pca=False, flat_hand_mean=True): hands default to zero (identity)."""
        N = betas.shape[0]
        z45 = torch.zeros(N, 45, dtype=betas.dtype, device=betas.device)
        lh = z45 if left_hand_pose is None else left_hand_pose
        rh = z45 if right_hand_pose is None else right_hand_pose
        full_pose = torch.cat([global_orient, body_pose, lh, rh], 1)
        verts, joints = self.lbs(betas, full_pose)
        joints = torch.cat([joints, verts[:, EXTRA_VERTEX_IDS]], 1)          # (N,73,3)
        if transl is not None:
            verts = verts + transl[:, None]
            joints = joints + transl[:, None]
        return verts, joints, full_pose


class OracleBodyModel(torch.nn.Module):
    """Duck-typed twin of the reference BodyModel (body_model.py:16-115) over SMPLHOracle."""

    def __init__(self, asset, num_betas=16, batch_size=1, use_vtx_selector=False,
                 dtype=torch.float32, device='cpu'):
        super().__init__()
        self.core = SMPLHOracle(asset, num_betas, dtype, device)
        self.model_type = 'smplh'
        self.num_joints = 51
        self.use_vtx_selector = use_vtx_selector
        self.batch_size = batch_size

    def forward(self, root_orient=None, pose_body=None, pose_hand=None, betas=None, trans=None, **kw):
        lh = rh = None
        if pose_hand is not None:
            lh, rh = pose_hand[:, :45], pose_hand[:, 45:]
        v, J, full_pose = self.core.forward(betas, root_orient, pose_body, trans, lh, rh)
        if not self.use_vtx_selector:
            J = J[:, :self.num_joints + 1]

        class _S:
            pass
        o = _S()
        o.v, o.Jtr, o.f, o.betas, o.pose_body, o.full_pose = v, J, self.core.faces, betas, pose_body, full_pose
        o.pose_hand = full_pose[:, 66:]
        return o
