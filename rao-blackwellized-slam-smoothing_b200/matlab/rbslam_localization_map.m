function m = rbslam_localization_map(m, foo, dVarft, sigma2)
%RBSLAM_LOCALIZATION_MAP  Attach the fixed GP map of the localisation example to a dense-mag model:
% foo (posterior-mean basis weights), dVarft (predictive variances read per particle) and sigma2, the
% constants captured by measModel in examples/mag-localization-mapping/run_localization.m:259-270.
  desc = rbslam_resolve(m.dynModel, m.measModel);
  desc.foo = double(foo(:)); desc.dVarft = double(dVarft); desc.sigma2 = double(sigma2);
  m.dynModel  = @(varargin) rbslam_handle_stub_(desc);
  m.measModel = @(varargin) rbslam_handle_stub_(desc);
end
function rbslam_handle_stub_(desc) %#ok<INUSD>
  error('rbslam:unsupportedModel', 'rbslam model handles are evaluated on the GPU and cannot be called on the host');
end
