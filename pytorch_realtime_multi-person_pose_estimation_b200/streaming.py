"""Pipelined capture -> infer -> draw loop for video files and cameras.

The reference's front-ends (/root/reference/video_demo.py:85-125, demo/web_demo.py:51-71) read a frame, call
get_outputs + paf_to_pose_cpp, draw, and only then read the next frame.  Here frames are grouped into batches and two
batches are kept in flight: while the GPU works on batch i+1 (upload, crop_with_factor, network, post-processing - the
fused engine), the host decodes the next frames and draws batch i.  The drawing itself is the reference's draw_humans.
"""
from .lib.utils.common import draw_humans
from .lib.utils.paf_to_pose import humans_from_rows
from . import _native as nat


def frames_of(capture, rotate_code=None, limit=None):
    """Frames of a cv2.VideoCapture-like object until it runs dry (optionally rotated, as video_demo.py:27-45 does)."""
    import cv2
    k = 0
    while limit is None or k < limit:
        ok, frame = capture.read()
        if not ok or frame is None:
            return
        if rotate_code is not None:
            frame = cv2.rotate(frame, rotate_code)
        yield frame
        k += 1


class PoseStream:
    """for frame, humans, drawn in PoseStream(model, frames, batch=8): ...

    model: a module from lib.network.rtpose_vgg.get_model() (bare or DataParallel-wrapped) on a CUDA device.
    frames: any iterator of uint8 BGR frames of ONE shape.  batch=1 minimises latency (camera), larger batches raise
    throughput (files).  `draw=False` skips the rendering."""

    def __init__(self, model, frames, batch=8, preprocess='rtpose', dest_size=368, factor=8, thresh=0.1, draw=True):
        core = getattr(model, "module", model)
        if not hasattr(core, "pose_engine"):
            raise nat.B200PoseError("PoseStream needs a model from lib.network.rtpose_vgg.get_model()")
        self.engine = core.pose_engine(batch_cap=max(1, int(batch)))
        self.engine.net.set_preprocess(preprocess)
        self.frames, self.batch = iter(frames), max(1, int(batch))
        self.dest_size, self.factor, self.thresh, self.draw = dest_size, factor, thresh, draw

    def _next_batch(self):
        out = []
        for frame in self.frames:
            out.append(frame)
            if len(out) == self.batch:
                break
        return out

    def _finish(self, ticket, batch):
        rows = self.engine.fetch_arrays(ticket=ticket)
        _, _, (ph, pw) = nat.crop_geometry(batch[0].shape[0], batch[0].shape[1], self.dest_size, self.factor)
        for frame, r in zip(batch, rows):
            humans = humans_from_rows(r, pw, ph)
            # draw_humans scales by the ORIGINAL frame size, like the reference front-ends (coordinates are normalised by
            # the padded network input, padding included - the same small offset the reference has)
            yield frame, humans, (draw_humans(frame, humans, imgcopy=True) if self.draw else None)

    def __iter__(self):
        pending = None
        while True:
            batch = self._next_batch()
            if batch:
                if any(f.shape != batch[0].shape for f in batch):
                    raise nat.B200PoseError("PoseStream: the frames of a stream must share one shape")
                ticket = self.engine.submit_images(batch, self.dest_size, self.factor, self.thresh)
            if pending is not None:
                yield from self._finish(*pending)
            if not batch:
                return
            pending = (ticket, batch)
