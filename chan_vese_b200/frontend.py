"""Python front-end with the reference's command line (src/main.cpp:583-1008) for hosts whose images are not PNM.

    python -m chan_vese_b200.frontend -i image.png -S -L 0.25 -T 100 -K 30 -N 70 -s -V -O

Same option table, defaults, validation messages and output naming as the reference (and as bin/chan_vese, the
C++14 front-end, which reads binary PNM only); decode / encode and the XVID video go through cv2, the one OpenCV
build this image has.  What the C++ front-end cannot do and this one can:

  * any image format cv2.imread knows (src/main.cpp:877-881);
  * -V: the per-step video "<stem>.avi" of the contour over the ORIGINAL image, frame 0 = the initial level set, one
    frame per CSV step, optional "t = n" overlay text (VideoWriterManager, src/VideoWriterManager.cpp:24-127; frames
    are fed by the asynchronous per-step mask observer of the C ABI, cvb_mask_fn);
  * --rect x,y,w,h / --circ cx,cy,r replace the interactive window of -R / -C (src/main.cpp:899-921).

All numerics run in the CUDA library through chan_vese_b200.Context; there is no CPU path here either.  `backend`
(tests only) is anything with perona_malik / csv_run / mask of the same signatures.
"""
import os
import sys

import numpy as np

# cv::Scalar(B, G, R) of ChanVese::Colors, src/main.cpp:110-117
COLORS = {
    "white": (255, 255, 255), "black": (0, 0, 0), "red": (0, 0, 255), "green": (0, 255, 0), "blue": (255, 0, 0),
    "magenta": (255, 0, 255), "yellow": (0, 255, 255), "cyan": (255, 255, 0),
}
TEXT_POSITIONS = {"TL": "TopLeft", "BL": "BottomLeft", "TR": "TopRight", "BR": "BottomRight"}


class MsgExit(Exception):
    """msg_exit(), src/main.cpp:173-178: the message goes to stderr framed by newlines, the exit status is failure."""


def add_suffix(path, suffix, delim="_"):
    """add_suffix(), src/main.cpp:158-167."""
    head, tail = os.path.split(path)
    stem, ext = os.path.splitext(tail)
    return os.path.join(head, stem + delim + suffix + ext)


def saturate_u8(u):
    """Mat::convertTo(CV_8UC1) of an fp64 matrix: round half to even, clamp to 0..255 (saturate_cast<uchar>)."""
    return np.clip(np.rint(u), 0, 255).astype(np.uint8)


class VideoWriterManager:
    """cv::VideoWriter wrapper of the reference (include/VideoWriterManager.hpp, src/VideoWriterManager.cpp)."""

    FONT_SCALE, FONT_THICKNESS = 0.8, 1  # FontParameters, src/FontParameters.cpp:5-11; face HERSHEY_PLAIN, type AA

    def __init__(self, input_filename, img, contour_color, fps, pos, enable_overlay, writer=None):
        import cv2
        self.cv2 = cv2
        self.img = img
        self.contour_color = tuple(int(c) for c in contour_color)
        self.pos = pos
        self.enable_overlay = enable_overlay
        self.frames = 0
        self.filename = os.path.splitext(input_filename)[0] + ".avi"  # change_extension(input, "avi"), :36
        h, w = img.shape[:2]
        self.vw = writer if writer is not None else cv2.VideoWriter(self.filename, cv2.VideoWriter_fourcc(*"XVID"), fps, (w, h))

    def draw_contour(self, dst, u):
        """:57-75.  Note the threshold: saturate_cast<uchar>(u) > 0, i.e. u > 0.5 -- not the u > 0 of separate()."""
        return self.draw_contour_mask(dst, (saturate_u8(u) > 0).astype(np.uint8))

    def draw_contour_mask(self, dst, mask):
        """The same from the 0/1 mask itself (the per-step mask observer of the C ABI delivers it: CVB_MASK_CONTOUR)."""
        cv2 = self.cv2
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        cs, hier = cv2.findContours(mask, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
        if hier is None:  # no contour at all: the reference would index an empty hierarchy
            return 0
        idx = 0
        while idx >= 0:  # the top-level contours; drawContours with the hierarchy also draws what they enclose
            cv2.drawContours(dst, cs, idx, self.contour_color, 1, 8, hier)
            idx = int(hier[0][idx][0])
        return 0

    def overlay_color(self, txt):
        """:77-127: the text anchor and black/white, whichever reads better on the patch under the text."""
        cv2 = self.cv2
        (tw, th), _ = cv2.getTextSize(txt, cv2.FONT_HERSHEY_PLAIN, self.FONT_SCALE, self.FONT_THICKNESS)
        rows, cols = self.img.shape[:2]
        pad = 5
        if self.pos == "TopLeft":
            p, q = (pad, pad + th), (pad, pad)
        elif self.pos == "TopRight":
            p, q = (cols - pad - tw, pad + th), (cols - pad - tw, pad)
        elif self.pos == "BottomLeft":
            p, q = (pad, rows - pad), (pad, rows - pad - th)
        else:
            p, q = (cols - pad - tw, rows - pad), (cols - pad - tw, rows - pad - th)
        x0, y0 = max(q[0], 0), max(q[1], 0)
        roi = self.img[y0:y0 + th, x0:x0 + tw]
        avgs = roi.reshape(-1, roi.shape[-1]).mean(axis=0) if roi.size else np.zeros(3)
        intensity = 0.114 * avgs[0] + 0.587 * avgs[1] + 0.299 * avgs[2]
        color = COLORS["black"] if 255 - intensity < 105 else COLORS["white"]
        return color, p

    def compose(self, u, overlay_text="", mask=None):
        frame = self.img.copy()
        if mask is not None:
            self.draw_contour_mask(frame, mask)
        else:
            self.draw_contour(frame, u)
        if self.enable_overlay:
            color, p = self.overlay_color(overlay_text)
            self.cv2.putText(frame, overlay_text, p, self.cv2.FONT_HERSHEY_PLAIN, self.FONT_SCALE, color, self.FONT_THICKNESS,
                             self.cv2.LINE_AA)
        return frame

    def write_frame(self, u, overlay_text=""):
        """:41-54."""
        self.vw.write(self.compose(u, overlay_text))
        self.frames += 1

    def write_mask_frame(self, mask, overlay_text=""):
        """write_frame for a consumer that received the thresholded level set instead of the level set."""
        self.vw.write(self.compose(None, overlay_text, mask=mask))
        self.frames += 1

    def release(self):
        if hasattr(self.vw, "release"):
            self.vw.release()


def _parser():
    import argparse

    class Parser(argparse.ArgumentParser):
        def error(self, message):  # std::exception from the parser -> "error: " + what(), src/main.cpp:871-874
            raise MsgExit("error: " + message)

    ap = Parser(prog="chan_vese", add_help=False, allow_abbrev=False)
    ap.add_argument("-h", "--help", action="store_true", help="this message")
    ap.add_argument("-i", "--input", help="input image")
    ap.add_argument("--mu", type=float, default=0.5, help="length penalty parameter (must be positive or zero)")
    ap.add_argument("--nu", type=float, default=0.0, help="area penalty parameter")
    ap.add_argument("--dt", type=float, default=1.0, help="timestep")
    ap.add_argument("--lambda1", type=float, nargs="+", help="penalty of variance inside the contour (default: 1's)")
    ap.add_argument("--lambda2", type=float, nargs="+", help="penalty of variance outside the contour (default: 1's)")
    ap.add_argument("-e", "--epsilon", type=float, default=1.0, help="smoothing parameter in Heaviside/delta")
    ap.add_argument("-t", "--tolerance", type=float, default=0.001, help="tolerance in stopping condition")
    ap.add_argument("-N", "--max-steps", type=int, default=-1, help="maximum nof iterations (negative means unlimited)")
    ap.add_argument("-f", "--fps", type=float, default=10.0, help="video fps")
    ap.add_argument("-P", "--overlay-pos", default="TL", help="overlay tex position; allowed only: TL, BL, TR, BR")
    ap.add_argument("-l", "--line-color", default="blue", help="contour color (allowed only: black, white, R, G, B, Y, M, C")
    ap.add_argument("-K", "--edge-coef", type=float, default=10.0, help="coefficient for enhancing edge detection in Perona-Malik")
    ap.add_argument("-L", "--laplacian-coef", type=float, default=0.25,
                    help="coefficient in the gradient FD scheme of Perona-Malik (must be [0, 1/4])")
    ap.add_argument("-T", "--segment-time", type=float, default=20.0, help="number of smoothing steps in Perona-Malik")
    ap.add_argument("-S", "--segment", action="store_true", help="segment the image with Perona-Malik beforehand")
    ap.add_argument("-g", "--grayscale", action="store_true", help="read in as grayscale")
    ap.add_argument("-V", "--video", action="store_true", help="enable video output (changes the extension to '.avi')")
    ap.add_argument("-O", "--overlay-text", action="store_true", help="add overlay text")
    ap.add_argument("-I", "--invert-selection", action="store_true", help="invert selected region (see: select)")
    ap.add_argument("-s", "--select", action="store_true",
                    help="separate the region encolosed by the contour (adds suffix '_selection')")
    ap.add_argument("-R", "--rectangle", action="store_true", help="rectangular contour; give it with --rect x,y,w,h")
    ap.add_argument("-C", "--circle", action="store_true", help="circular contour; give it with --circ cx,cy,r")
    ap.add_argument("--rect", help="rectangular initial contour x,y,w,h (implies -R)")
    ap.add_argument("--circ", help="circular initial contour cx,cy,r (implies -C)")
    return ap


def _ints(opt, text, n):
    try:
        vals = [int(v) for v in text.split(",")]
    except ValueError:
        vals = []
    if len(vals) != n:
        raise MsgExit("error: option '--%s' needs %d comma-separated integers" % (opt, n))
    return vals


def _validate(o):
    """src/main.cpp:786-869, message for message."""
    if o.input is None:
        raise MsgExit("Error: you have to specify input file name!")
    if not os.path.exists(o.input):
        raise MsgExit('Error: file "%s" does not exists!' % o.input)
    if o.dt <= 0:
        raise MsgExit("Cannot have negative or zero timestep: %f." % o.dt)
    if o.mu < 0:
        raise MsgExit("Length penalty parameter cannot be negative: %f." % o.mu)
    if o.epsilon <= 0:
        # deliberate deviation: the reference's own check (src/main.cpp:831-834) tests an option name that does not exist
        # and never fires; epsilon = 0 divides by zero in the regularised delta.  Same message as bin/chan_vese (cli/).
        raise MsgExit("Cannot have negative or zero smoothing parameter: %f." % o.epsilon)
    n = 1 if o.grayscale else 3
    for name in ("lambda1", "lambda2"):
        lam = getattr(o, name)
        if lam is None:
            setattr(o, name, [1.0] * n)
            continue
        if o.grayscale and len(lam) != 1:
            raise MsgExit("Too many %s values for a grayscale image." % name)
        if not o.grayscale and len(lam) != 3:
            raise MsgExit("Number of %s values must be 3 for a colored input image." % name)
        if any(v < 0 for v in lam):
            raise MsgExit(("The value of %s cannot be negative." if o.grayscale else "Any value of %s cannot be negative.") % name)
    if o.overlay_pos.upper() not in TEXT_POSITIONS:
        raise MsgExit("Invalid text position requested.\nCorrect values are: TL -- top left\n"
                      "                    BL -- bottom left\n                    TR -- top right\n"
                      "                    BR -- bottom right")
    if o.line_color.lower() not in COLORS:
        raise MsgExit("Invalid contour color requested.\nCorrect values are: red, green, blue, black, white, magenta, yellow, cyan.")
    if o.laplacian_coef > 0.25 or o.laplacian_coef < 0:
        raise MsgExit("The Laplacian coefficient in Perona-Malik segmentation must be between 0 and 0.25.")
    if o.segment_time < o.laplacian_coef:
        raise MsgExit("The segmentation duration must exceed the value of Laplacian coefficient, %f." % o.laplacian_coef)
    if o.rect:
        o.rect = _ints("rect", o.rect, 4)
        o.rectangle = True
    if o.circ:
        o.circ = _ints("circ", o.circ, 3)
        o.circle = True
    if o.rectangle and o.circle:
        raise MsgExit("Cannot initialize with both rectangular and circular contour")
    if o.rectangle and not o.rect:
        raise MsgExit("No GUI on this host: give the rectangle as --rect x,y,w,h.")
    if o.circle and not o.circ:
        raise MsgExit("No GUI on this host: give the circle as --circ cx,cy,r.")
    if (o.rect and (o.rect[2] <= 0 or o.rect[3] <= 0)) or (o.circ and o.circ[2] <= 0):
        raise MsgExit("You must specify the contour with non-zero dimensions")


def run(argv, backend=None, video_writer=None):
    """main(), src/main.cpp:583-1008.  Returns the process exit status; stdout stays silent (except --help)."""
    import cv2

    import chan_vese_b200 as cv

    ap = _parser()
    o = ap.parse_args(argv)
    if o.help:
        print(ap.format_help())
        return 0
    _validate(o)
    # :877-886: grayscale is read as one plane and converted back to three for the outputs that carry colour
    raw = cv2.imread(o.input, cv2.IMREAD_GRAYSCALE if o.grayscale else cv2.IMREAD_COLOR)
    if raw is None:
        raise MsgExit('Error on opening "%s" (probably not an image)!' % o.input)
    img = cv2.cvtColor(raw, cv2.COLOR_GRAY2RGB) if o.grayscale else raw
    h, w = img.shape[:2]
    nch = 1 if o.grayscale else 3
    if o.rectangle:
        u = cv.levelset_rect(h, w, *o.rect)
    elif o.circle:
        u = cv.levelset_circ(h, w, *o.circ)
    else:
        u = cv.levelset_checkerboard(h, w)
    own = backend is None
    if own:
        backend = cv.Context(0)  # raises without a GPU: there is no CPU path
    try:
        vwm = None
        if o.video:  # :926-931
            vwm = VideoWriterManager(o.input, img, COLORS[o.line_color.lower()], o.fps, TEXT_POSITIONS[o.overlay_pos.upper()],
                                     o.overlay_text, writer=video_writer)
            vwm.write_frame(u, "t = 0" if o.overlay_text else "")
        channels = [np.ascontiguousarray(img[..., k]) for k in range(nch)]  # cv::split order, :933-937
        if o.segment:  # :939-947
            channels, _ = backend.perona_malik(channels, o.edge_coef, o.laplacian_coef, o.segment_time)
            smoothed = channels[0] if nch == 1 else np.stack(channels, axis=-1)
            cv2.imwrite(add_suffix(o.input, "pm"), smoothed)
        frame = None
        if vwm is not None:
            def frame(uu, step):  # :997
                vwm.write_frame(uu, ("t = %d" % step) if o.overlay_text else "")
                return 0
        params = cv.make_params(o.mu, o.nu, o.dt, o.epsilon, o.lambda1, o.lambda2, nch=nch)
        if vwm is not None and hasattr(backend, "csv_run_masks"):
            # the CUDA backend streams the per-step contour masks (1/64 of the bytes of u) through a pinned ring while
            # later steps run; the frames are the same as those drawn from u (same threshold, VideoWriterManager.cpp:65-68)
            def on_mask(m, step):
                vwm.write_mask_frame(m, ("t = %d" % step) if o.overlay_text else "")
                return 0
            u, steps, norm = backend.csv_run_masks(channels, u, params, on_mask, o.tolerance, o.max_steps, contour_rule=True)
        else:
            u, steps, norm = backend.csv_run(channels, u, params, o.tolerance, o.max_steps, frame)
        if vwm is not None:
            vwm.release()
        if o.select:  # :1004-1005, separate() :386-405
            m = np.asarray(backend.mask(u, o.invert_selection)).astype(bool)
            sel = np.full_like(img, 255)
            sel[m] = img[m]
            cv2.imwrite(add_suffix(o.input, "selection"), sel)
    finally:
        if own:
            backend.close()
    return 0


def main(argv=None):
    try:
        return run(sys.argv[1:] if argv is None else argv)
    except MsgExit as e:
        sys.stderr.write("\n%s\n\n" % e)
        return 1


if __name__ == "__main__":
    sys.exit(main())
