function [U, nablaU] = interpolate_U(background_flow1, background_flow2, alpha, x, h)
% [U, nablaU] = interpolate_U(bf1, bf2, alpha, x, h)   -- qg_flow_ray_trace/interpolate_U.m
% The two frames are uploaded once per (bf1, bf2) pair and blended on the device.  qgsw_raytrace.m passes NEW frames every
% flow step (:141-143) and the same pair at every ode23 stage within it (:261): the pair is recognised by comparing the
% full contents of all twelve arrays with the copies kept here (copy-on-write, so free), never by a checksum.
persistent eng c1 c2 cpar
nx = size(background_flow1.u, 1);
par = [nx, h];
if isempty(eng) || ~isequal(par, cpar) || ~isequal(background_flow1, c1) || ~isequal(background_flow2, c2)
    if isempty(eng) || ~isequal(par, cpar)
        if ~isempty(eng), swrt_mex('destroy', eng); eng = []; end
        eng = swrt_mex('create', nx, h * nx, 1, 1, 1, 0, 1e-10);       % mode 1 = LAGRANGE6, device 0, bump 1e-10 (the
                                                                       % interpolate.m beside interpolate_U.m, qg_flow_ray_trace/interpolate.m:13)
    end
    b = background_flow1; swrt_mex('set_flow_grid', eng, 0, b.u, b.v, b.ux, b.uy, b.vx, b.vy);
    b = background_flow2; swrt_mex('set_flow_grid', eng, 1, b.u, b.v, b.ux, b.uy, b.vx, b.vy);
    c1 = background_flow1; c2 = background_flow2; cpar = par;
end
[u, v, ux, uy, vx, vy] = swrt_mex('eval_at', eng, alpha, x(:, 1), x(:, 2));
U = [u, v];
nablaU.u_x = ux; nablaU.u_y = uy; nablaU.v_x = vx; nablaU.v_y = vy;
end
