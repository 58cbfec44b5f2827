"""Drop-in for path_plan/hybrid_a_star.py: Node, hybrid_a_star (hybrid_a_star.py:22-389).

The step-wise methods keep the reference's signatures.  Their data-parallel parts run on the GPU
(successor rollout + sub-step collision checks + rs lengths: avp_expand_pure; rs shot: avp_rs_optimal
+ avp_collision_check; Dijkstra: avp_dijkstra_query); the bookkeeping (Python heapq open list,
closed list, node objects) stays in Python exactly as in the reference.  Whole searches should use
PathPlanner.a_star_plan, which runs everything on the device."""
import math
import queue

import numpy as np

from ..map.costmap import Map, Vehicle
from ..collision_check import collision_check
from .compute_h import Dijkstra
from . import rs_curve


class Node:
    def __init__(self, index=None, x=0.0, y=0.0, theta=0.0, parent_index=None, child_index=None, is_in_openlist=False,
                 is_in_closedlist=False, is_forward=True, steering_angle=None) -> None:
        self.index = index
        self.x = x
        self.y = y
        self.theta = theta
        self.parent_index = parent_index
        self.child_index = child_index
        self.in_open = is_in_openlist
        self.in_closed = is_in_closedlist
        self.forward = is_forward
        self.steering_angle = steering_angle
        self.h = 0
        self.g = 0
        self.f = 0

    def __lt__(self, other):
        return bool(self.f < other.f)


class hybrid_a_star:
    def __init__(self, config: dict, park_map: Map, vehicle: Vehicle) -> None:
        self.vehicle = vehicle
        self.steering_angle = np.linspace(-vehicle.max_steering_angle, vehicle.max_steering_angle, config['steering_angle_num'])
        self.park_map = park_map
        self.heuristic = Dijkstra(park_map)
        _, self.h_value_list = self.heuristic.compute_path(node_x=park_map.case.x0, node_y=park_map.case.y0)
        self.global_index = 0
        self.config = config
        self.open_list = queue.PriorityQueue()
        self.closed_list = []
        self.dt = config['dt']
        self.ddt = config['trajectory_dt']
        self.initial_node = Node(x=park_map.case.x0, y=park_map.case.y0, index=0, theta=rs_curve.pi_2_pi(park_map.case.theta0))
        self.goal_node = Node(x=park_map.case.xf, y=park_map.case.yf, theta=rs_curve.pi_2_pi(park_map.case.thetaf))
        self.open_list.put(self.initial_node)
        self.initial_node.in_open = True
        self.max_delta_heading = vehicle.max_v * np.tan(vehicle.max_steering_angle) / vehicle.lw * self.dt
        if config['collision_check'] == 'circle':
            self.collision_checker = collision_check.two_circle_checker(vehicle=vehicle, map=park_map, config=config)
        else:
            self.collision_checker = collision_check.distance_checker(vehicle=vehicle, map=park_map, config=config)

    # -- hybrid_a_star.py:126-241
    def expand_node(self, current_node: Node) -> queue.PriorityQueue:
        child_group = queue.PriorityQueue()
        n = int(2 * self.config['steering_angle_num'])
        dev = self.collision_checker._device()
        poses, flags, rs_len = dev.expand_pure(0, [current_node.x, current_node.y, current_node.theta])   # rollout, checks, rs.L on the GPU
        b = self.park_map.boundary
        for i in range(n):
            steering_angle = self.steering_angle[i % self.config['steering_angle_num']]
            is_forward = i < n / 2
            x_, y_, theta_ = np.float64(poses[i, 0]), np.float64(poses[i, 1]), np.float64(poses[i, 2])
            skip = False
            for c in self.closed_list:
                if c.x == x_ and c.y == y_ and c.theta == theta_:
                    skip = True
                    break
                elif x_ > b[1] or x_ < b[0] or y_ > b[3] or y_ < b[2]:
                    skip = True
                    break
            if skip:
                continue
            child_node = None
            for o in self.open_list.queue:
                if o.x == x_ and o.y == y_ and o.theta == theta_:
                    child_node = o
            if child_node is None:
                child_node = Node(x=x_, y=y_, theta=theta_, index=self.global_index + i + 1, parent_index=current_node.index,
                                  is_forward=is_forward, steering_angle=steering_angle)
                if flags[i] & 1:
                    self.closed_list.append(child_node)
                    child_node.in_closed = True
                else:
                    child_node.g = self.calc_node_cost(child_node, father_theta=current_node.theta, father_gear=current_node.forward)
                    child_node.h = self.calc_node_heuristic(child_node, rs_length=rs_len[i])
                    child_node.f = child_node.g + child_node.h
                    self.open_list.put(child_node)
                    child_node.in_open = True
            else:
                new_h = self.calc_node_heuristic(child_node, rs_length=rs_len[i])
                new_g = self.calc_node_cost(child_node, father_theta=current_node.theta, father_gear=current_node.forward)
                new_f = new_h + new_g
                if new_f < child_node.f:
                    child_node.f, child_node.g, child_node.h = new_f, new_g, new_h
                    child_node.parent_index = current_node.index
                    child_node.forward = is_forward
                    child_node.steering_angle = steering_angle
            if child_node.in_closed is False and child_node.in_open is True:
                child_group.put(child_node)
        current_node.in_closed = True
        current_node.in_open = False
        self.closed_list.append(current_node)
        self.global_index += n
        return child_group

    # -- hybrid_a_star.py:243-259
    def calc_node_cost(self, node: Node, father_theta, father_gear):
        cost_gear = self.config['cost_gear'] if node.forward != father_gear else 0
        cost = cost_gear + self.config['cost_heading_change'] * abs(node.theta - father_theta)
        return self.config['cost_scale'] * cost

    # -- hybrid_a_star.py:261-298
    def calc_node_heuristic(self, current_node: Node, rs_length=None):
        gid = self.park_map.convert_position_to_index(grid_x=current_node.x, grid_y=current_node.y)
        h1 = self.h_value_list.lookup(gid)
        if h1 is None:
            h1, self.h_value_list = self.heuristic.compute_path(node_x=current_node.x, node_y=current_node.y)
        if rs_length is None:
            max_c = 1 / self.vehicle.min_radius_turn
            rs_length = rs_curve.calc_optimal_path(current_node.x, current_node.y, current_node.theta, self.goal_node.x,
                                                   self.goal_node.y, self.goal_node.theta, max_c).L
        return max(h1 / 100, rs_length)

    # -- hybrid_a_star.py:300-349
    def try_reach_goal(self, current_node: Node):
        collision, rs_path, in_radius, collision_p = False, None, False, None
        distance = np.sqrt((current_node.x - self.goal_node.x) ** 2 + (current_node.y - self.goal_node.y) ** 2)
        if distance < self.config['flag_radius']:
            in_radius = True
            rs_path, collision, collision_p = self.try_rs_curve(current_node)
        return rs_path, collision, {'in_radius': in_radius, 'collision_position': collision_p}

    def try_rs_curve(self, current_node: Node):
        max_c = 1 / self.vehicle.min_radius_turn
        rs_path = rs_curve.calc_optimal_path(current_node.x, current_node.y, current_node.theta, self.goal_node.x, self.goal_node.y,
                                             self.goal_node.theta, max_c)
        poses = [[x, y, rs_curve.pi_2_pi(t)] for x, y, t in zip(rs_path.x, rs_path.y, rs_path.yaw)]
        hits = self.collision_checker.check_many(poses)          # all course points in one launch
        idx = np.nonzero(hits)[0]
        if len(idx):
            return rs_path, True, poses[int(idx[0])]
        return rs_path, False, None

    # -- hybrid_a_star.py:351-389
    def finish_path(self, current_node: Node):
        node = current_node
        chain = []
        while node.index != 0:
            chain.append(node)
            for c in self.closed_list:
                if c.index == node.parent_index:
                    node = c
                    break
        chain.append(node)
        all_path = [[node.x, node.y, node.theta]]
        for i in range(len(chain)):
            k = len(chain) - 1 - i
            if k == 0:
                break
            for j in range(math.ceil(self.dt / self.ddt)):
                speed = self.vehicle.max_v if chain[k - 1].forward else -self.vehicle.max_v
                td_j = speed * self.ddt * (j + 1)
                theta_j = chain[k].theta + (self.vehicle.max_v * np.tan(chain[k - 1].steering_angle)) / self.vehicle.lw * self.ddt * (j + 1)
                theta_j = rs_curve.pi_2_pi(theta_j)
                all_path.append([chain[k].x + td_j * np.cos(theta_j), chain[k].y + td_j * np.sin(theta_j), theta_j])
        return all_path
