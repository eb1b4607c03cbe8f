"""GOP structure recovered from the video header: which frame is I / P / hierarchical B,
its references and the coding <-> display order maps.

Host-side mirror of the reference's ``coolchic/utils/codingstructure.py`` (``Frame`` :23-155,
``CodingStructure`` :158-436, ``compute_coding_struct`` :267-436, accessors :603-647): same
class / method names and the same resulting order, but built recursively (each gap between
two coded frames is bisected depth-first) instead of by repeated list scans.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, List, Optional


@dataclass
class Frame:
    coding_order: int
    display_order: int
    depth: int = 0
    index_references: List[int] = field(default_factory=list)
    data: Optional[Any] = None  # FrameData once decoded
    frame_type: str = field(init=False)

    def __post_init__(self):
        assert len(self.index_references) <= 2, (
            "A frame can not have more than 2 references.\n"
            f"Found {len(self.index_references)} references for frame {self.display_order} "
            "(display order).\n Exiting!"
        )
        self.index_references.sort()
        self.frame_type = "IPB"[len(self.index_references)]

    def set_frame_data(self, data: Any) -> None:
        self.data = data


@dataclass
class CodingStructure:
    n_frames: int = 1
    intra_pos: List[int] = field(default_factory=lambda: [0])
    p_pos: List[int] = field(default_factory=list)
    frames: List[Frame] = field(init=False)

    def __post_init__(self):
        self.intra_pos = sorted(self.intra_pos)
        self.p_pos = sorted(self.p_pos)
        assert self.intra_pos and self.intra_pos[0] == 0, (
            "First frame of the video should an intra frame. Change --intra_pos to include the frame 0."
        )
        last_ok = self.intra_pos[-1] == self.n_frames - 1 or (
            bool(self.p_pos) and self.p_pos[-1] == self.n_frames - 1
        )
        assert last_ok, (
            "Last frame of the video should be either an intra frame or a P frame. "
            "Add -1 to --intra_pos or --p_pos to include the last frame."
        )
        common = sorted(set(self.intra_pos).intersection(self.p_pos))
        assert not common, (
            "Frames can not be an I-frame and a P-frame at the same time!\n"
            f"Found --intra_pos={self.intra_pos} --p_pos={self.p_pos}.\n"
            f"Frame(s) {common} are in both arguments, they should be present only in one of them."
        )
        self.frames = self._build()

    def _build(self) -> List[Frame]:
        """Reference algorithm (codingstructure.py:267-436): 1) all I frames in list order,
        2) P frames, each referencing the closest already-placed frame in the past,
        3) while frames are missing: take the first uncoded display index, bisect the gap
        between its closest coded neighbours.  Step 3 visits gaps left to right and, inside
        a gap, always descends into the left half first."""
        by_disp = {}
        order: List[Frame] = []

        def add(disp: int, refs: List[int], depth: int) -> None:
            if disp in by_disp:  # duplicates in the lists are tolerated by the reference
                return
            f = Frame(coding_order=len(order), display_order=disp, depth=depth, index_references=refs)
            by_disp[disp] = f
            order.append(f)

        for d in self.intra_pos:
            add(d, [], 0)
        for d in self.p_pos:
            past = max((x for x in by_disp if x < d), default=min(by_disp))
            add(d, [past], by_disp[past].depth + 1)

        def fill(lo: int, hi: int) -> None:
            if hi - lo < 2:
                return
            mid = lo + (hi - lo) // 2
            add(mid, [lo, hi], max(by_disp[lo].depth, by_disp[hi].depth) + 1)
            fill(lo, mid)
            fill(mid, hi)

        anchors = sorted(by_disp)
        for lo, hi in zip(anchors[:-1], anchors[1:]):
            fill(lo, hi)
        assert len(order) == self.n_frames, f"coding structure has {len(order)} frames, expected {self.n_frames}"
        return sorted(order, key=lambda f: f.display_order)

    # ---- accessors (codingstructure.py:603-647) ---------------------------------------
    def get_max_depth(self) -> int:
        return max(f.depth for f in self.frames)

    def get_all_frames_of_depth(self, depth: int) -> List[Frame]:
        return [f for f in self.frames if f.depth == depth]

    def get_max_coding_order(self) -> int:
        return max(f.coding_order for f in self.frames)

    def get_max_display_order(self) -> int:
        return max(f.display_order for f in self.frames)

    def get_frame_from_coding_order(self, coding_order: int) -> Optional[Frame]:
        for f in self.frames:
            if f.coding_order == coding_order:
                return f
        return None

    def get_frame_from_display_order(self, display_order: int) -> Optional[Frame]:
        for f in self.frames:
            if f.display_order == display_order:
                return f
        return None

    def pretty_structure_diagram(self) -> str:
        """One line per temporal layer, e.g. ``I0 P8`` / ``B4`` / ``B2 B6`` ..."""
        if self.n_frames == 1:
            return "I0"
        lines = []
        for depth in range(self.get_max_depth() + 1):
            lines.append(
                " ".join(f"{f.frame_type}{f.display_order}" for f in self.get_all_frames_of_depth(depth))
            )
        return "\n".join(lines)
