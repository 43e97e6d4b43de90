-- | Drop-in for "Data.Text.AhoCorasick.Splitter" (reference: src/Data/Text/AhoCorasick/Splitter.hs:13-22): a
-- single-needle automaton; splitting is the reference's own fold (`stepAccum` :158-170, `finalizeAccum` :140-147) over
-- the matches the device returns.  NOT COMPILED HERE -- see INTEGRATION.md.
module Data.Text.AhoCorasick.Splitter
  ( Splitter, automaton, build, separator, split, splitIgnoreCase, splitReverse, splitReverseIgnoreCase
  ) where

import Data.Function (on)
import Data.List.NonEmpty (NonEmpty ((:|)))
import Data.Text.Utf8 (CodeUnitIndex (..), Text)

import qualified Data.List.NonEmpty as NonEmpty
import qualified Data.Text as Text
import qualified Data.Text.Utf8 as Utf8

import Data.Text.AhoCorasick.Automaton (AcMachine)
import qualified Data.Text.AhoCorasick.Automaton as Aho

data Splitter = Splitter
  { splitterAutomaton :: AcMachine ()    -- INVARIANT: exactly one needle (:48-52)
  , splitterSeparator :: Text
  }

instance Eq Splitter where (==) = (==) `on` separator                                 -- :175-177

build :: Text -> Splitter                                                             -- :63-67
build sep = Splitter (Aho.build [(sep, ())]) sep

automaton :: Splitter -> AcMachine ()
automaton = splitterAutomaton

separator :: Splitter -> Text
separator = splitterSeparator

split, splitIgnoreCase :: Splitter -> Text -> NonEmpty Text                           -- :84-85, :95-96
split = (NonEmpty.reverse .) . splitReverse
splitIgnoreCase = (NonEmpty.reverse .) . splitReverseIgnoreCase

-- | :98-106.  The accumulator is (fragments so far, start of the current fragment).
splitReverse :: Splitter -> Text -> NonEmpty Text
splitReverse s t = finalizeAccum t $ Aho.runText zeroAccum (stepAccum sepLength t) (automaton s) t
  where sepLength newFragmentStart = newFragmentStart - Utf8.lengthUtf8 (separator s)

-- | :109-118: the separator must be lower case; its length in the text is found by walking code points back.
splitReverseIgnoreCase :: Splitter -> Text -> NonEmpty Text
splitReverseIgnoreCase s t = finalizeAccum t $ Aho.runLower zeroAccum (stepAccum sepStart t) (automaton s) t
  where sepStart newFragmentStart = Utf8.skipCodePointsBackwards t (newFragmentStart - 1) (Text.length (separator s) - 1)

data Accum = Accum ![Text] !CodeUnitIndex
zeroAccum :: Accum
zeroAccum = Accum [] 0                                                                -- :150-152

-- | :158-170: a separator that starts before the current fragment overlaps the previous one and is ignored.
stepAccum :: (CodeUnitIndex -> CodeUnitIndex) -> Text -> Accum -> Aho.Match () -> Aho.Next Accum
stepAccum sepStartOf hay acc@(Accum res fragmentStart) (Aho.Match newFragmentStart _)
  | sepStart < fragmentStart = Aho.Step acc
  | otherwise = Aho.Step (Accum (Utf8.unsafeSliceUtf8 fragmentStart (sepStart - fragmentStart) hay : res) newFragmentStart)
  where sepStart = sepStartOf newFragmentStart

finalizeAccum :: Text -> Accum -> NonEmpty Text                                       -- :140-147
finalizeAccum hay (Accum res fragmentStart) = Utf8.unsafeSliceUtf8 fragmentStart (Utf8.lengthUtf8 hay - fragmentStart) hay :| res
