"""Drop-in for gym_rrt/envs/grid_cell_rrt.py (/root/reference/gym_rrt/envs/grid_cell_rrt.py:36-90): one
square of the planner's grid with its heading subsections.  The node lists live on the GPU; a
subsection's `node_array` is a read-only view fetched from the planner on access."""
import numpy as np


def angle_wrap(ang):
    """(-pi, pi] wrap by repeated +-2pi, as grid_cell_rrt.py:13-30"""
    while not (-np.pi <= ang <= np.pi):
        ang += (-2 * np.pi) if ang > np.pi else (2 * np.pi)
    return ang


class Grid_cell_RRT:
    def __init__(self, x, y, side_length=1, num_of_subsections=8, planner=None, flat_base=0):
        self.x = x
        self.y = y
        self.side_length = side_length
        self.delta_theta = float(2.0 * np.pi) / float(num_of_subsections)
        self.subsection_cells = []
        theta = 0.0
        for i in range(num_of_subsections):
            self.subsection_cells.append(self.Subsection_grid_cell_RRT(theta, planner, flat_base + i))
            theta = angle_wrap(theta + self.delta_theta)

    def has_node(self):
        return any(len(s.node_array) != 0 for s in self.subsection_cells)

    def __repr__(self):
        return "RRT Grid: [x=%s, y=%s, side length=%s], node list: %s" % (self.x, self.y, self.side_length,
                                                                         self.subsection_cells)

    __str__ = __repr__

    class Subsection_grid_cell_RRT:
        def __init__(self, theta, planner=None, flat_id=0):
            self.theta = theta
            self.flat_id = flat_id          # (row * cols + col) * subsections + subsection
            self._planner = planner

        @property
        def node_array(self):
            return [] if self._planner is None else self._planner._nodes_in(self.flat_id)

        def __repr__(self):
            return "Subsec: theta=%s, node list: %s" % (self.theta, self.node_array)

        __str__ = __repr__
