function desc = rbslam_resolve(varargin)
%RBSLAM_RESOLVE  Map dynModel/measModel/dynResNorm handles to one model descriptor.
% Any handle that was not made by rbslam_model raises rbslam:unsupportedModel --
% there is no CPU fallback by design.
  desc = [];
  for k = 1:numel(varargin)
    fh = varargin{k};
    if isempty(fh), continue; end
    ok = false;
    if isa(fh, 'function_handle')
      info = functions(fh);
      if isfield(info, 'workspace') && ~isempty(info.workspace) && isfield(info.workspace{1}, 'desc')
        d = info.workspace{1}.desc; ok = true;
      end
    end
    if ~ok
      error('rbslam:unsupportedModel', ['handle %d is not a registered rbslam model handle ' ...
            '(denseMag3D, denseRadio2D, sparseVisual2D); arbitrary closures cannot run on the GPU'], k);
    end
    if isempty(desc), desc = d; elseif ~isequal(desc, d)
      error('rbslam:unsupportedModel', 'handles belong to different models');
    end
  end
end
