"""Glyph atlas for the label text of `DrawEffect` (watsor/output/draw.py:86-88: `cv2.putText(..., FONT_HERSHEY_DUPLEX,
font_scale 0.5, white, thickness 1, cv2.LINE_AA)`).

The reference's text IS OpenCV's rasteriser: Hershey strokes in 16.16 fixed point, anti-aliased lines blended into the
8-bit image one stroke after the other.  Neither the stroke tables nor the line code are restated here.  The atlas is
built BY CONSTRUCTION from the installed OpenCV, once per process, on the host:

* OpenCV draws glyph after glyph; glyph *i* starts at pen position `org.x + sum(advance_j, j < i) / 2` pixels (advances
  are integers in half pixels at scale 0.5), so a glyph's rendering depends only on the glyph, the half-pixel PHASE of
  its pen position and -- when the text runs over the right image border, where OpenCV clips and re-caps the strokes --
  the distance from the pen to the border.  It does not depend on the integer pen position, on the row, or on the
  other glyphs.
* whatever sequence of blends a glyph applies to a pixel, the result is a function of the pixel's previous value only
  (channels are independent, the colour is white in every channel): a 256-entry table per sprite pixel.

So the atlas holds, per (glyph, phase, clip distance), a `[rows][cols][256]` uint8 table obtained by rendering that
glyph over uniform backgrounds of every level, and the GPU kernel applies the tables glyph by glyph in text order --
which reproduces `cv2.putText` bit for bit (tests/test_effects_host.py checks that on the CPU with a numpy emulation
of the kernel, tests/test_gpu_effects.py on the GPU).  Text metrics (`cv2.getTextSize`, draw.py:55-59) follow from the
same advances.
"""
import numpy as np

FONT_SCALE = 0.5          # coco.py:118
FONT_THICKNESS = 1        # coco.py:117
BASE_ORG = (8, 24)        # where probes are drawn inside their scratch image
PHASE1_PREFIX = 'i '      # odd total advance (9 + 16 half pixels) and no ink near the probed glyph


def _cv2():
    import cv2
    return cv2


class FontAtlas:
    """advance[c] in half pixels, text height / baseline, sprite window and the table
    `lut[glyph][phase][clip][row][col][level]` (clip index k-1 for a border k = 1..cols pixels right of the pen,
    index cols for "further away": unclipped)."""

    def __init__(self, charset):
        cv2 = _cv2()
        self.font = cv2.FONT_HERSHEY_DUPLEX
        chars = sorted(set(charset) | set(PHASE1_PREFIX))
        assert all(32 <= ord(c) < 127 for c in chars), 'printable ASCII only (cv2.putText of Hershey fonts)'
        self.chars = chars
        self.index = {c: i for i, c in enumerate(chars)}
        # advance: getTextSize(cc) = cvRound(2 * adv / 2 + thickness) = adv + 1   (adv in half pixels)
        self.advance = {c: cv2.getTextSize(c + c, self.font, FONT_SCALE, FONT_THICKNESS)[0][0] - 1 for c in chars}
        (_, self.text_height), self.baseline = cv2.getTextSize('x', self.font, FONT_SCALE, FONT_THICKNESS)
        assert sum(self.advance[c] for c in PHASE1_PREFIX) % 2 == 1
        self._probe_window()
        self._build()

    # ---- metrics (draw.py:55-59)
    def text_width(self, text):
        total = sum(self.advance[c] for c in text)
        return int(np.rint(total * 0.5 + FONT_THICKNESS))      # cvRound: half to even

    # ---- construction
    def _render(self, text, width, height, levels):
        """putText over a 4-channel image whose channels are uniform at `levels` (4 backgrounds per call)."""
        cv2 = _cv2()
        img = np.empty((height, width, 4), np.uint8)
        img[:] = np.asarray(levels, np.uint8)
        cv2.putText(img, text, BASE_ORG, self.font, FONT_SCALE, (255, 255, 255, 255), FONT_THICKNESS, cv2.LINE_AA)
        return img

    def _pen(self, phase):
        half = sum(self.advance[c] for c in PHASE1_PREFIX) if phase else 0
        return (PHASE1_PREFIX if phase else ''), BASE_ORG[0] + half // 2

    def _probe_window(self):
        """Ink extent over all glyphs and phases (relative to the integer pen position and the text origin row)."""
        x_lo = y_lo = 10 ** 6
        x_hi = y_hi = -10 ** 6
        for c in self.chars:
            for phase in (0, 1):
                prefix, px = self._pen(phase)
                a = self._render(prefix + c, px + 64, 64, (0, 0, 0, 0))[:, :, 0]
                b = self._render(prefix, px + 64, 64, (0, 0, 0, 0))[:, :, 0] if prefix else np.zeros_like(a)
                ys, xs = np.nonzero(a != b)
                if len(xs):
                    x_lo, x_hi = min(x_lo, xs.min() - px), max(x_hi, xs.max() - px)
                    y_lo, y_hi = min(y_lo, ys.min() - BASE_ORG[1]), max(y_hi, ys.max() - BASE_ORG[1])
        assert x_lo >= 0, 'a glyph inks left of its pen position: the left border would clip it'
        self.x0, self.y0 = 0, int(y_lo)
        self.cols, self.rows = int(x_hi) + 1, int(y_hi - y_lo) + 1

    def _build(self):
        n, rows, cols = len(self.chars), self.rows, self.cols
        nclip = cols + 1
        self.lut = np.empty((n, 2, nclip, rows, cols, 256), np.uint8)
        height = BASE_ORG[1] + self.y0 + rows + 8
        ident = np.arange(256, dtype=np.uint8)
        for gi, c in enumerate(self.chars):
            for phase in (0, 1):
                prefix, px = self._pen(phase)
                for k in range(1, nclip + 1):
                    width = px + (k if k <= cols else cols + 16)
                    dst = self.lut[gi, phase, k - 1]
                    dst[:] = ident                                  # columns beyond the border stay untouched
                    w = min(cols, width - px)
                    y = BASE_ORG[1] + self.y0
                    for lv in range(0, 256, 4):
                        img = self._render(prefix + c, width, height, (lv, lv + 1, lv + 2, lv + 3))
                        dst[:, :w, lv:lv + 4] = img[y:y + rows, px:px + w, :]
        self.lut.setflags(write=False)

    # ---- numpy emulation of the kernel's text loop (tests, and the definition of what the kernel must do)
    def draw(self, image, text, org):
        """In place; image HxWxC uint8 (every channel gets white text); returns nothing."""
        h, w = image.shape[:2]
        pen2 = 0
        for ch in text:
            px = org[0] + (pen2 >> 1)
            k = w - px
            if k >= 1:
                clip = min(k, self.cols + 1) - 1
                t = self.lut[self.index[ch], pen2 & 1, clip]
                for r in range(self.rows):
                    y = org[1] + self.y0 + r
                    if 0 <= y < h:
                        for col in range(min(self.cols, k)):
                            x = px + col
                            if x >= 0:
                                image[y, x] = t[r, col][image[y, x]]
            pen2 += self.advance[ch]
