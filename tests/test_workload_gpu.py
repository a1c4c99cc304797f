"""The bench workload's step driver (arfe_b200/workload.py): plain launches with the RoI
plan on a second stream, and the same step replayed as one CUDA graph, give the same bits."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _outputs(st):
    return [t.clone() for t in st.dx + st.dy + st.y + [st.dbsf, st.z, st.d_ori, st.d_ab, st.gathered]]


def test_step_graph_replay_matches_plain_launches(cuda):
    from arfe_b200 import workload as wl
    host = wl.host_inputs(batch=2, rois_per_img=40, channels=256, img_h=160, img_w=224, channels_last=True)
    st = wl.TrainStep(host, cuda)
    st.step()
    st.step()
    torch.cuda.synchronize()
    ref = _outputs(st)
    assert all(torch.isfinite(t.float()).all() for t in ref)
    st.capture()
    assert st.graph is not None
    for t in st.dx + st.dy + [st.dbsf, st.z]:
        t.zero_()
    st.step()  # one graph replay
    st.step()
    torch.cuda.synchronize()
    for a, b in zip(ref, _outputs(st)):
        assert torch.equal(a, b)
    assert st.launches_per_step() == 14
