// ref_glue_plan.cpp -- flat C entry points around the REFERENCE's own FootPlanner and FootTrajectoryManager
// (TEST INFRASTRUCTURE ONLY).  oracle/ref_build.sh compiles src/quadruped_controller/foot_planner.cpp and
// trajectory.cpp from where they lie under /root/reference against the stand-in headers of oracle/ref_stubs and links
// this file.  Pins: FootPlanner::singleFoot, FootTrajectory::initSystem / constantTerms / trackTrajectory,
// FootTrajectoryManager's phase -> t mapping and centre point.  Not pinned: arma::solve (stand-in: LU with partial
// pivoting, as LAPACK dgesv).  The planner's stance->swing bookkeeping (FootPlanner::updateStates) is exercised
// through positions() with a fresh planner per call, which plans exactly the legs that are in swing.
#include <map>
#include <string>
#include <vector>

#include <quadruped_controller/foot_planner.hpp>
#include <quadruped_controller/trajectory.hpp>

#include "plan_oracle.h"

using namespace quadruped_controller;

namespace
{
const char* kLegNames[4] = { "RL", "FL", "RR", "FR" };

arma::mat33 to_mat33(const double* R)
{
  arma::mat33 m;
  for (unsigned i = 0; i < 3; i++)
    for (unsigned j = 0; j < 3; j++) m(i, j) = R[3 * i + j];
  return m;
}
arma::vec3 to_vec3(const double* v) { return { v[0], v[1], v[2] }; }
}  // namespace

extern "C" {

// FootPlanner::singleFoot (foot_planner.cpp:76-104); the planner's constants are hard-coded in its constructor
void ref_single_foot(double t_stance, int leg, const orc_state* s, double out[3])
{
  const FootPlanner planner;
  const arma::vec3 f = planner.singleFoot(t_stance, to_mat33(s->Rwb), to_vec3(s->x), to_vec3(s->xdot), to_vec3(s->w),
                                          to_vec3(s->xdot_d), to_vec3(&s->feet[3 * leg]), kLegNames[leg]);
  for (int i = 0; i < 3; i++) out[i] = f(i);
}

// one tick of commander_node.cpp:429-461 for one robot, through the reference's own classes
void ref_plan(const orc_plan_params* p, const orc_state* s, orc_plan_rec* plan, orc_swing* sw)
{
  GaitMap gait_map;
  FootholdMap foot_actual_map;
  for (int leg = 0; leg < 4; leg++)
  {
    gait_map.emplace(kLegNames[leg], std::make_pair(s->contact[leg] ? LegState::stance : LegState::swing, plan->phase[leg]));
    foot_actual_map.emplace(kLegNames[leg], to_vec3(&s->feet[3 * leg]));
  }
  const arma::mat33 Rwb = to_mat33(s->Rwb);
  const arma::vec3 x = to_vec3(s->x);
  // a fresh planner plans every leg that is in swing (updateStates with an empty state map, foot_planner.cpp:113-133)
  const FootPlanner planner;
  const auto plan_result = planner.positions(p->t_stance, Rwb, x, to_vec3(s->xdot), to_vec3(s->w), to_vec3(s->xdot_d),
                                             foot_actual_map, gait_map);
  const FootholdMap planned = std::get<FootholdMap>(plan_result);
  FootTrajBoundsMap bounds;
  for (int leg = 0; leg < 4; leg++)
  {
    if (s->contact[leg]) continue;
    if (plan->replan[leg])
    {
      const arma::vec3 p_final = planned.at(kLegNames[leg]);
      const arma::vec3 p_start = Rwb * foot_actual_map.at(kLegNames[leg]) + x;  // commander_node.cpp:452
      for (int i = 0; i < 3; i++)
      {
        plan->p_final[3 * leg + i] = p_final(i);
        plan->p_start[3 * leg + i] = p_start(i);
      }
      plan->replan[leg] = 0;
    }
    bounds.emplace(kLegNames[leg], FootTrajBounds(to_vec3(&plan->p_start[3 * leg]), to_vec3(&plan->p_final[3 * leg])));
  }
  const FootTrajectoryManager manager(p->height, p->t_swing, p->t_stance);
  const FootStateMap states = manager.referenceStates(gait_map, bounds);
  for (int leg = 0; leg < 4; leg++)
  {
    const auto it = states.find(kLegNames[leg]);
    if (it == states.end()) continue;
    for (int i = 0; i < 3; i++)
    {
      sw->foot_ref_pos[3 * leg + i] = it->second.position(i);
      sw->foot_ref_vel[3 * leg + i] = it->second.velocity(i);
    }
  }
}

// Two ticks through ONE FootTrajectoryManager, i.e. with the state the reference keeps between ticks (traj_map_):
// tick 1 plans leg a; tick 2 plans leg b while leg a is mid-swing.  referenceStates(gait_map, bounds) clears traj_map_
// before it adds the newly planned legs (trajectory.cpp:316), so after tick 2 the reference no longer has a trajectory
// for leg a and referenceState(a) returns a zero FootState (trajectory.cpp:366-388).  out_a / out_b: position (3) +
// velocity (3) of the two legs after tick 2; found_a: whether leg a still had a trajectory.
void ref_two_tick_replan(const orc_plan_params* p, int leg_a, const double a_start[3], const double a_final[3], double phase_a,
                         int leg_b, const double b_start[3], const double b_final[3], double phase_b, double out_a[6],
                         double out_b[6])
{
  const FootTrajectoryManager manager(p->height, p->t_swing, p->t_stance);
  GaitMap gait_map;
  for (int leg = 0; leg < 4; leg++)
  {
    const bool swing = leg == leg_a || leg == leg_b;
    gait_map.emplace(kLegNames[leg], std::make_pair(swing ? LegState::swing : LegState::stance,
                                                    leg == leg_a ? phase_a : (leg == leg_b ? phase_b : 0.0)));
  }
  FootTrajBoundsMap first, second;
  first.emplace(kLegNames[leg_a], FootTrajBounds(to_vec3(a_start), to_vec3(a_final)));
  second.emplace(kLegNames[leg_b], FootTrajBounds(to_vec3(b_start), to_vec3(b_final)));
  GaitMap tick1 = gait_map;
  tick1.at(kLegNames[leg_b]).first = LegState::stance;  // leg b lifts off one tick later
  manager.referenceStates(tick1, first);
  manager.referenceStates(gait_map, second);
  const FootState sa = manager.referenceState(kLegNames[leg_a], phase_a);  // commander_node.cpp:487-488
  const FootState sb = manager.referenceState(kLegNames[leg_b], phase_b);
  for (int i = 0; i < 3; i++)
  {
    out_a[i] = sa.position(i);
    out_a[3 + i] = sa.velocity(i);
    out_b[i] = sb.position(i);
    out_b[3 + i] = sb.velocity(i);
  }
}

void ref_plan_batch(const orc_plan_params* p, const orc_state* s, orc_plan_rec* plan, orc_swing* sw, long long n)
{
  for (long long i = 0; i < n; i++) ref_plan(p, &s[i], &plan[i], &sw[i]);
}

}  // extern "C"
