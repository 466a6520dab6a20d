//! `extern "C"` bindings of include/prl.h (ABI version 2). One declaration per exported symbol;
//! tests/test_abi.py checks that libprl.so exports every symbol named here.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const PRL_OK: c_int = 0;
pub const PRL_ERR_NO_GOAL: c_int = -4;
pub const PRL_ERR_CUDA: c_int = -5;
pub const PRL_ERR_OVERFLOW: c_int = -6;

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct prl_object {
    pub position: [f32; 3],
    pub scale: [f32; 2],
    pub rotation: f32,
    /// 0 = Block{fixed:true}, 1 = Block{fixed:false}, 2 = Goal (common.rs:46-49)
    pub kind: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct prl_ga_config {
    pub number_of_steps: i32,
    pub number_of_agents: i64,
    pub repeat_move: i32,
    pub mutation_rate: f32,
    pub keep_best: i32,
    pub seed: u64,
}

// Layout of the two PODs that cross the ABI by value/pointer: sizes and field offsets of the C
// structs in include/prl.h (static_asserted there too, csrc/prl_world.cc; tests/test_abi.py
// compares these numbers with the ctypes mirror).
pub const PRL_ABI_VERSION: i32 = 2;
pub const PRL_OBJECT_SIZE: usize = 28;
pub const PRL_GA_CONFIG_SIZE: usize = 40;
const _: () = {
    use std::mem::{align_of, size_of};
    assert!(size_of::<prl_object>() == PRL_OBJECT_SIZE && align_of::<prl_object>() == 4);
    assert!(size_of::<prl_ga_config>() == PRL_GA_CONFIG_SIZE && align_of::<prl_ga_config>() == 8);
};
#[cfg(test)]
mod layout {
    use super::*;
    use std::mem::offset_of;
    #[test]
    fn field_offsets_match_include_prl_h() {
        assert_eq!((offset_of!(prl_object, position), offset_of!(prl_object, scale), offset_of!(prl_object, rotation), offset_of!(prl_object, kind)), (0, 12, 20, 24));
        assert_eq!(
            (offset_of!(prl_ga_config, number_of_steps), offset_of!(prl_ga_config, number_of_agents), offset_of!(prl_ga_config, repeat_move),
             offset_of!(prl_ga_config, mutation_rate), offset_of!(prl_ga_config, keep_best), offset_of!(prl_ga_config, seed)),
            (0, 8, 16, 20, 24, 32)
        );
        assert_eq!(unsafe { prl_abi_version() }, PRL_ABI_VERSION);
    }
}

pub enum prl_world {}
pub enum prl_batch {}
pub enum prl_ga {}

#[link(name = "prl")]
extern "C" {
    pub fn prl_last_error() -> *const c_char;
    pub fn prl_abi_version() -> i32;
    pub fn prl_device_count() -> i32;
    pub fn prl_debug_bounds_violations() -> i32;

    pub fn prl_world_create(player_position: *const f32, objs: *const prl_object, n_objects: i32, out: *mut *mut prl_world) -> c_int;
    pub fn prl_world_from_json(utf8: *const c_char, len: usize, out: *mut *mut prl_world) -> c_int;
    pub fn prl_world_to_json(w: *const prl_world, buf: *mut c_char, cap: usize, needed: *mut usize) -> c_int;
    pub fn prl_world_num_objects(w: *const prl_world) -> i32;
    pub fn prl_world_get_object(w: *const prl_world, index: i32, out: *mut prl_object) -> c_int;
    pub fn prl_world_player_position(w: *const prl_world, out: *mut f32) -> c_int;
    pub fn prl_world_counts(w: *const prl_world, n_bodies: *mut i32, n_colliders: *mut i32, n_goals: *mut i32, n_pairs: *mut i32) -> c_int;
    pub fn prl_world_object_body(w: *const prl_world, index: i32) -> i32;
    pub fn prl_world_destroy(w: *mut prl_world);

    pub fn prl_batch_create(w: *const prl_world, n_envs: i64, device: i32, stream: *mut c_void, out: *mut *mut prl_batch) -> c_int;
    pub fn prl_batch_destroy(b: *mut prl_batch);
    pub fn prl_batch_num_envs(b: *const prl_batch) -> i64;
    pub fn prl_batch_num_bodies(b: *const prl_batch) -> i32;
    pub fn prl_batch_reset(b: *mut prl_batch) -> c_int;
    pub fn prl_batch_step(b: *mut prl_batch, moves: *const u8) -> c_int;
    pub fn prl_batch_step_uniform(b: *mut prl_batch, mv: u32, n_steps: i32) -> c_int;
    pub fn prl_batch_distance_to_goals(b: *mut prl_batch, out: *mut f32) -> c_int;
    pub fn prl_batch_won(b: *mut prl_batch, out: *mut u8) -> c_int;
    pub fn prl_batch_get_state(b: *mut prl_batch, env0: i64, n: i64, out: *mut f32) -> c_int;
    pub fn prl_batch_rollout(b: *mut prl_batch, genomes: *const u8, genome_len: i32, repeat_move: i32, number_of_steps: i32,
                             early_exit: i32, out_score: *mut f32, out_step_of_win: *mut i32) -> c_int;
    pub fn prl_batch_rollout_info(b: *mut prl_batch, out_score: *mut f32, out_step_of_win: *mut i32, out_steps: *mut i32) -> c_int;
    pub fn prl_batch_trajectory(b: *mut prl_batch, n: i64, genomes: *const u8, genome_len: i32, repeat_move: i32, number_of_steps: i32,
                                out_state: *mut f32, out_distance: *mut f32, out_won: *mut u8) -> c_int;
    pub fn prl_batch_observation_dim(b: *const prl_batch) -> i32;
    pub fn prl_batch_get_observation(b: *mut prl_batch, out: *mut f32) -> c_int;
    pub fn prl_policy_num_params(b: *const prl_batch, hidden: i32) -> i32;
    pub fn prl_batch_rollout_policy(b: *mut prl_batch, params: *const f32, hidden: i32, number_of_steps: i32, early_exit: i32,
                                    out_score: *mut f32, out_step_of_win: *mut i32) -> c_int;
    pub fn prl_batch_sync(b: *mut prl_batch) -> c_int;
    pub fn prl_batch_launch_count(b: *const prl_batch) -> i64;
    pub fn prl_batch_state_bytes(b: *const prl_batch) -> i64;

    pub fn prl_ga_create(w: *const prl_world, cfg: *const prl_ga_config, device: i32, stream: *mut c_void, rank: i32, n_ranks: i32,
                         out: *mut *mut prl_ga) -> c_int;
    pub fn prl_ga_create_multi(w: *const prl_world, cfg: *const prl_ga_config, devices: *const i32, n_devices: i32, out: *mut *mut prl_ga) -> c_int;
    pub fn prl_ga_destroy(g: *mut prl_ga);
    pub fn prl_ga_num_shards(g: *const prl_ga) -> i32;
    pub fn prl_ga_init(g: *mut prl_ga) -> c_int;
    pub fn prl_ga_buffers(g: *mut prl_ga, local_scores: *mut *mut f32, local_genomes: *mut *mut u8, global_scores: *mut *mut f32,
                          global_genomes: *mut *mut u8, local_count: *mut i64, local_offset: *mut i64, genome_len: *mut i32) -> c_int;
    pub fn prl_ga_best(g: *mut prl_ga, best_score: *mut f32, best_genome: *mut u8, best_index: *mut i64) -> c_int;
    pub fn prl_ga_next_generation(g: *mut prl_ga) -> c_int;
    pub fn prl_ga_step_generation(g: *mut prl_ga, best_score: *mut f32, best_genome: *mut u8, best_index: *mut i64) -> c_int;
    pub fn prl_ga_read_generation(g: *mut prl_ga, scores: *mut f32, genomes: *mut u8) -> c_int;
    pub fn prl_ga_generation(g: *const prl_ga) -> i64;
    pub fn prl_ga_launch_count(g: *const prl_ga) -> i64;
    pub fn prl_ga_env_steps(g: *mut prl_ga) -> i64;
}
